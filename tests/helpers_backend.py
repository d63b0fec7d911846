"""Test doubles for the backend tests: a deterministic fake embedding provider/router (the dict-shaped
provider result of SURVEY Appendix A.8) and an oracle-backed engine (test infrastructure only)."""
import hashlib
import json
import os
from types import SimpleNamespace

import numpy as np

from oracle import retrieval_oracle as O


def text_vec(text, dim):
    seed = int.from_bytes(hashlib.sha256(text.encode()).digest()[:8], "little")
    v = np.random.default_rng(seed).standard_normal(dim).astype(np.float32)
    return v / np.linalg.norm(v)


class FakeRouter:
    def __init__(self, dim, as_dict=True):
        self.dim, self.as_dict = dim, as_dict
        self.policy = {"embedding_provider": "fake"}
        self.providers = {"fake": SimpleNamespace(kwargs={"embed_model": "fake-embed"})}
        self.calls = []

    def embed(self, model_hint, texts, require=None):
        self.calls.append((model_hint, list(texts), require))
        vecs = [text_vec(t, self.dim).tolist() for t in texts]
        return {"vectors": vecs} if self.as_dict else vecs


class ListSink:
    def __init__(self):
        self.events = []

    def record(self, ev):
        self.events.append(ev)


class OracleEngine:
    """Same interface as backend.CudaEngine, computed with the oracle on bf16-rounded inputs."""

    def __init__(self, X_f32, row_ptr=None, col_idx=None, metric="cos"):
        self.X = O.bf16_bits_to_f32(O.f32_to_bf16_bits(X_f32))
        self.N, self.D = self.X.shape
        self.row_ptr, self.col_idx, self.metric = row_ptr, col_idx, metric

    def _q(self, qvecs):
        return O.bf16_bits_to_f32(O.f32_to_bf16_bits(np.asarray(qvecs, dtype=np.float32)))

    def dense_search(self, qvecs, k):
        s, i = O.topk(self._q(qvecs), self.X, k, metric=self.metric)
        return s.tolist(), i.tolist()

    def has_graph(self):
        return self.row_ptr is not None

    def expand_rescore(self, qvecs, seeds, window, alpha_dense, alpha_graph, k):
        out = {key: [] for key in ("score", "id", "hop", "score_graph_norm", "score_dense_norm", "cos", "ncand", "status")}
        for q, sd in zip(self._q(qvecs), seeds):
            top, dist = O.expand_rescore(q, self.X, self.row_ptr, self.col_idx, sd, window, alpha_dense, alpha_graph, k)
            cos = {c: float(O.score_matrix(q[None], self.X[[c]], self.metric)[0, 0]) for c, *_ in top}
            pad = k - len(top)
            out["score"].append([t[1] for t in top] + [float("-inf")] * pad)
            out["id"].append([t[0] for t in top] + [-1] * pad)
            out["hop"].append([dist[t[0]] for t in top] + [-1] * pad)
            out["score_graph_norm"].append([t[3] for t in top] + [0.0] * pad)
            out["score_dense_norm"].append([t[4] for t in top] + [0.0] * pad)
            out["cos"].append([cos[t[0]] for t in top] + [0.0] * pad)
            out["ncand"].append(len(dist))
            out["status"].append(0)
        return out


def make_corpus(tmpdir, n_docs=12, sents_per_doc=5, dim=64):
    """HotpotQA-shaped docs.jsonl (schema of my_code/ingest_hotpotqa.py:73-81) + embeddings + CSR on disk."""
    from a_modular_rag_framework_b200 import store
    docs = []
    for d in range(n_docs):
        for s in range(sents_per_doc):
            docs.append({"doc_id": f"Title {d}#{s}", "title": f"Title {d}", "sent_id": s, "text": f"sentence {s} of document {d} about topic{d % 4}"})
    index_path = os.path.join(tmpdir, "docs.jsonl")
    with open(index_path, "w", encoding="utf-8") as f:
        for doc in docs:
            f.write(json.dumps(doc) + "\n")
    X = np.stack([text_vec(doc["text"], dim) for doc in docs])
    emb_path = os.path.join(tmpdir, "embeddings")
    store.save_embeddings(emb_path, O.f32_to_bf16_bits(X), "bf16")
    src, dst = store.next_in_doc_edges(docs)
    row_ptr, col_idx = store.csr_from_edges(len(docs), src, dst)
    graph_path = os.path.join(tmpdir, "graph_csr.npz")
    store.save_csr(graph_path, row_ptr, col_idx)
    return SimpleNamespace(docs=docs, X=X, index_path=index_path, emb_path=emb_path, graph_path=graph_path,
                           row_ptr=row_ptr, col_idx=col_idx, dim=dim)
