"""The backend on the CUDA engine: same flows as tests/test_backend_host.py, results compared with the oracle engine, the
reference's recorded BFS cases replayed through graph_root/<graph_id>/graph.json, batches, and several (virtual) devices."""
import json
import os

import numpy as np
import pytest

from a_modular_rag_framework_b200.backend import B200RetrievalBackend
from tests.helpers_backend import FakeRouter, OracleEngine, make_corpus, make_graph_id_fixture
from tests.test_backend_host import _req, check_bfs_case_through_backend

pytestmark = pytest.mark.gpu

TOL = 5e-4      # fused scores: 1e-3 relative on bf16 cosines through the min-max map (tests/helpers_graph.py derives the bound)


def _same_hits(a, b):
    sa, sb = [h["score"] for h in a["hits"]], [h["score"] for h in b["hits"]]
    assert len(sa) == len(sb) and np.allclose(sa, sb, rtol=0, atol=TOL)
    ra, rb = [h["meta"]["row"] for h in a["hits"]], [h["meta"]["row"] for h in b["hits"]]
    if ra != rb:   # only near-ties may swap
        for x, y, u, v in zip(ra, rb, sa, sb):
            assert x == y or abs(u - v) <= TOL
    for ha, hb in zip(a["hits"], b["hits"]):
        if ha["meta"]["row"] == hb["meta"]["row"]:
            assert ha["meta"].get("distance") == hb["meta"].get("distance")
            assert ha["id"] == hb["id"] and ha["meta"]["source"] == hb["meta"]["source"]


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
            _same_hits(a, b)
            assert a["hits"][0]["meta"]["row"] == row
            assert a["diagnostics"]["graph_candidates"] == b["diagnostics"]["graph_candidates"]
            assert a["diagnostics"]["engine"] == "CudaEngine"


def test_graph_id_graphs_replay_reference_bfs_cases_on_the_gpu(tmp_path, golden_dir):
    """graph_root/<graph_id>/graph.json -> cached local CSR -> q_match seeds -> bfs_kernel, through backend.run: the recorded
    outputs of the reference's expand_qmatch_neighbors (bfs_cases.json) come back exactly (candidates, hop distances)."""
    with open(os.path.join(golden_dir, "bfs_cases.json")) as f:
        cases = json.load(f)
    for n, case in enumerate(cases):
        fx = make_graph_id_fixture(str(tmp_path / f"c{n}"), case, "g")
        be = B200RetrievalBackend(router=FakeRouter(fx.dim), index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root)
        out = check_bfs_case_through_backend(be, case, "g")
        ref = B200RetrievalBackend(router=FakeRouter(fx.dim), index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root,
                                   engine=OracleEngine(fx.X))
        _same_hits(out, ref.run(_req("the question", graph_id="g", top_k=max(1, len(case["nodes"])), graph_window=case["window"])))


def test_retrieve_many_on_the_gpu_equals_one_at_a_time(tmp_path, golden_dir):
    """f-4 on the CUDA engine: a batch with different graph_ids, a corpus-level graph request, a dense-only request and an
    unusable query -> ONE dense search + one expansion per graph kind; same hits as retrieve() per request and as the oracle."""
    with open(os.path.join(golden_dir, "bfs_cases.json")) as f:
        cases = json.load(f)
    fx = make_graph_id_fixture(str(tmp_path), cases[7], "ga")
    g2 = {"nodes": [], "edges": [{"source": f"Doc::sent{v}", "target": f"Doc::sent{v + 1}", "type": "next_in_doc"} for v in range(0, 59, 2)]
          + [{"source": "q1", "target": "Doc::sent10", "type": "q_match"}, {"source": "q1", "target": "Doc::sent41", "type": "q_match"}]}
    os.makedirs(os.path.join(fx.graph_root, "gb"))
    with open(os.path.join(fx.graph_root, "gb", "graph.json"), "w") as f:
        json.dump(g2, f)
    kw = dict(index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root)
    be = B200RetrievalBackend(router=FakeRouter(fx.dim), **kw)
    ref = B200RetrievalBackend(router=FakeRouter(fx.dim), engine=OracleEngine(fx.X), **kw)
    reqs = [_req("the question", graph_id="ga", top_k=20), _req("alpha3 sentence", graph_id="gb", top_k=20),
            _req(fx.docs[33]["text"], graph_id="ga", top_k=7), _req(fx.docs[5]["text"], graph_id="", top_k=5),
            _req("x", graph_id="nope", top_k=5)] * 3
    many = be.retrieve_many(reqs)
    for r, m in zip(reqs, many):
        one = be.retrieve(r)
        assert [h.id for h in one.hits] == [h.id for h in m.hits]
        assert [h.score for h in one.hits] == [h.score for h in m.hits]
        o = ref.run(r)
        _same_hits({"hits": [{"id": h.id, "score": h.score, "meta": h.meta} for h in m.hits]}, o)
    assert {h.meta["row"] for h in many[1].hits} == {10, 11, 40, 41}
    # columnar access: rows / fused scores without materialising hits
    batch = be.run_batch(reqs[:4])
    assert batch.rows(3)[0] == 5 and len(batch.rows(2)) == 7 and np.all(np.diff(batch.scores(0)) <= 0)


def test_several_devices_behind_the_plugin(tmp_path):
    """B200RetrievalBackend(devices=[...]): one process drives several row shards through the C-ABI exchange.  Two VIRTUAL
    devices (the same GPU twice) exercise the whole path on a 1-GPU box: dense search, graph stage, hit materialisation."""
    c = make_corpus(str(tmp_path), n_docs=50, sents_per_doc=6, dim=64)
    kw = dict(router=FakeRouter(c.dim), index_path=c.index_path, emb_path=c.emb_path, graph_path=c.graph_path, seed_k=6)
    one = B200RetrievalBackend(**kw)
    for devs in ([0, 0], [0, 0, 0]):
        multi = B200RetrievalBackend(devices=devs, **kw)
        assert len(multi.engine.shards) == len(devs) and multi.engine.comms is not None
        reqs = [_req(c.docs[row]["text"], graph_id=g, top_k=10) for row in (0, 17, 101, 299) for g in ("", "g")]
        for a, b in zip(one.retrieve_many(reqs), multi.retrieve_many(reqs)):
            assert [h.id for h in a.hits] == [h.id for h in b.hits]
            assert np.allclose([h.score for h in a.hits], [h.score for h in b.hits], rtol=0, atol=1e-6)
            assert a.diagnostics["graph_candidates"] == b.diagnostics["graph_candidates"]
