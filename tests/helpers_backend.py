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
    """Same interface as backend.CudaEngine (arrays in, arrays out), computed with the oracle on bf16-rounded inputs."""

    def __init__(self, X_f32, row_ptr=None, col_idx=None, metric="cos"):
        self.X = O.bf16_bits_to_f32(O.f32_to_bf16_bits(X_f32))
        self.N, self.D = self.X.shape
        self.row_ptr, self.col_idx, self.metric = row_ptr, col_idx, metric
        self.calls = {"dense": 0, "expand": 0}

    def _q(self, Q):
        return O.bf16_bits_to_f32(O.f32_to_bf16_bits(np.asarray(Q, dtype=np.float32)))

    def dense_search(self, Q, k):
        self.calls["dense"] += 1
        s, i = O.topk(self._q(Q), self.X, k, metric=self.metric)
        pad = k - s.shape[1]
        return (np.pad(s.astype(np.float32), ((0, 0), (0, pad)), constant_values=-np.inf),
                np.pad(i, ((0, 0), (0, pad)), constant_values=-1))

    def has_graph(self):
        return self.row_ptr is not None

    def mmr(self, ids, scores, top_k, lam):
        out = np.full((len(ids), top_k), -1, dtype=np.int32)
        for b in range(len(ids)):
            items = [(j, float(scores[b][j]), self.X[int(r)].astype(float).tolist() if 0 <= int(r) < self.N else None)
                     for j, r in enumerate(ids[b]) if r >= 0]
            picks = O.mmr_diversify(items, top_k=top_k, lambda_weight=lam)
            out[b, :len(picks)] = [p[0] for p in picks]
        return out

    def expand_rescore(self, Q, seeds, window, alpha_dense, alpha_graph, k, graph=None):
        self.calls["expand"] += 1
        row_ptr, col_idx, node_rows = (self.row_ptr, self.col_idx, None) if graph is None else graph
        B = len(seeds)
        out = {"score": np.full((B, k), -np.inf, np.float32), "id": np.full((B, k), -1, np.int64), "hop": np.full((B, k), -1, np.int32),
               "score_graph_norm": np.zeros((B, k), np.float32), "score_dense_norm": np.zeros((B, k), np.float32),
               "cos": np.zeros((B, k), np.float32), "ncand": np.zeros(B, np.int32), "status": np.zeros(B, np.int32)}
        for b, (q, sd) in enumerate(zip(self._q(Q), seeds)):
            top, dist = O.expand_rescore(q, self.X, row_ptr, col_idx, sd, window, alpha_dense, alpha_graph, k, node_rows=node_rows)
            n = len(top)
            ids = [t[0] for t in top]
            out["score"][b, :n] = [t[1] for t in top]
            out["id"][b, :n] = ids
            out["hop"][b, :n] = [dist[i] for i in ids]
            out["score_graph_norm"][b, :n] = [t[3] for t in top]
            out["score_dense_norm"][b, :n] = [t[4] for t in top]
            out["cos"][b, :n] = [float(O.score_matrix(q[None], self.X[[i]], self.metric)[0, 0]) if 0 <= i < self.N else 0.0 for i in ids]
            out["ncand"][b] = len(dist)
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


def make_graph_id_fixture(tmpdir, case, gid, dim=32, qmatch=True, extra_edges=True):
    """A corpus of case['V'] single-document sentences + data/graph/<gid>/graph.json shaped like the reference's
    (GraphConstructionNetworkX output read by load_graph_json, graph_utils.py:13-22): sentence nodes
    '<title>::sent<j>', next_in_doc edges case['src'] -> case['dst'], q_match edges q1 -> seeds, plus edge / node
    types build_index ignores.  -> namespace(index_path, emb_path, graph_root, docs, X)"""
    from a_modular_rag_framework_b200 import store
    V = case["V"]
    docs = [{"doc_id": f"Doc#{v}", "title": "Doc", "sent_id": v, "text": f"alpha{v} beta{v % 7} sentence"} for v in range(V)]
    os.makedirs(tmpdir, exist_ok=True)
    index_path = os.path.join(tmpdir, "docs.jsonl")
    with open(index_path, "w", encoding="utf-8") as f:
        for d in docs:
            f.write(json.dumps(d) + "\n")
    X = np.stack([text_vec(d["text"], dim) for d in docs])
    emb_path = os.path.join(tmpdir, "embeddings")
    store.save_embeddings(emb_path, O.f32_to_bf16_bits(X), "bf16")
    nodes = [{"id": "q1", "type": "question", "props": {"text": "the question"}}]
    nodes += [{"id": f"Doc::sent{v}", "type": "sentence", "props": {"text": docs[v]["text"], "doc": "Doc"}} for v in range(V)]
    nodes += [{"id": "doc::Doc", "type": "document", "props": {}}, {"id": "ent::Alpha", "type": "entity", "props": {}}]
    edges = [{"source": f"Doc::sent{a}", "target": f"Doc::sent{b}", "type": "next_in_doc"} for a, b in zip(case["src"], case["dst"])]
    if qmatch:
        edges += [{"source": "q1", "target": f"Doc::sent{s}", "type": "q_match"} for s in case["seeds"]]
    if extra_edges:
        edges += [{"source": f"Doc::sent{v}", "target": "doc::Doc", "type": "in_doc"} for v in range(min(V, 5))]
        edges += [{"source": "Doc::sent0", "target": f"Doc::sent{V - 1}", "type": "semantic_sim"},
                  {"source": "q2", "target": "Doc::sent1", "type": "q_match"},
                  {"source": "Doc::sent0", "target": "ent::Alpha", "type": "mentions"}]
    graph_root = os.path.join(tmpdir, "graph")
    os.makedirs(os.path.join(graph_root, gid), exist_ok=True)
    with open(os.path.join(graph_root, gid, "graph.json"), "w", encoding="utf-8") as f:
        json.dump({"nodes": nodes, "edges": edges}, f)
    return SimpleNamespace(index_path=index_path, emb_path=emb_path, graph_root=graph_root, docs=docs, X=X, dim=dim)
