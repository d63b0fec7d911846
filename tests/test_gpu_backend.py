"""The backend on the CUDA engine: same flow as tests/test_backend_host.py, results compared with the oracle engine."""
import numpy as np
import pytest

from a_modular_rag_framework_b200.backend import B200RetrievalBackend
from tests.helpers_backend import FakeRouter, OracleEngine, make_corpus
from tests.test_backend_host import _req

pytestmark = pytest.mark.gpu


def test_cuda_engine_matches_oracle_engine(tmp_path):
    c = make_corpus(str(tmp_path), n_docs=40, sents_per_doc=6, dim=64)
    cuda = B200RetrievalBackend(router=FakeRouter(c.dim), index_path=c.index_path, emb_path=c.emb_path, graph_path=c.graph_path, seed_k=6)
    ref = B200RetrievalBackend(router=FakeRouter(c.dim), index_path=c.index_path, emb_path=c.emb_path, graph_path=c.graph_path, seed_k=6,
                               engine=OracleEngine(c.X, c.row_ptr, c.col_idx))
    assert type(cuda.engine).__name__ == "CudaEngine"
    for row in (0, 17, 101, 239):
        for graph_id in ("g", ""):
            a = cuda.run(_req(c.docs[row]["text"], graph_id=graph_id, top_k=10))
            b = ref.run(_req(c.docs[row]["text"], graph_id=graph_id, top_k=10))
            sa, sb = [h["score"] for h in a["hits"]], [h["score"] for h in b["hits"]]
            assert np.allclose(sa, sb, atol=2e-3)
            ra, rb = [h["meta"]["row"] for h in a["hits"]], [h["meta"]["row"] for h in b["hits"]]
            if ra != rb:   # only near-ties may swap
                for x, y, u, v in zip(ra, rb, sa, sb):
                    assert x == y or abs(u - v) <= 2e-3
            assert a["hits"][0]["meta"]["row"] == row
            assert a["diagnostics"]["graph_candidates"] == b["diagnostics"]["graph_candidates"]
            assert a["diagnostics"]["engine"] == "CudaEngine"
