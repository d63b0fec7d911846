#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN FUNCTIONS.

Run once in the build container (where /root/reference is mounted, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests or golden vectors, so these fixtures are what pins the
oracle (oracle/retrieval_oracle.py) and, through it, the CUDA path.  Every expected value below is
produced by code imported from /root/reference:

* DenseReranker._cosine                 app/modules/retrieval/retrieval_backend.py:192-197
* HybridRetrievalBackend._minmax_norm   app/modules/retrieval/retrieval_backend.py:296-301
* build_index / expand_qmatch_neighbors app/modules/retrieval/graph_utils.py:25-55, 58-129
* mmr_diversify                         app/utils/similarity.py:33-62 (own seeded stream, appended last so that the
  earlier fixtures stay byte-identical)
* the stable ``sort(key=score, reverse=True)[:k]`` rule of retrieval_backend.py:371-372 is applied
  here with the same Python ``list.sort`` the reference uses, on candidates enumerated in id order.

The GPU box has no /root/reference; tests only read the committed outputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("B200R_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from app.modules.retrieval.retrieval_backend import DenseReranker, HybridRetrievalBackend  # noqa: E402
from app.modules.retrieval.graph_utils import build_index, expand_qmatch_neighbors  # noqa: E402

from oracle.retrieval_oracle import f32_to_bf16_bits, bf16_bits_to_f32  # noqa: E402  (dtype helpers only)


def ref_cosine(a, b):
    return DenseReranker._cosine(None, list(a), list(b))


def ref_minmax(values):
    return HybridRetrievalBackend._minmax_norm(None, dict(values))


def ref_sorted_topk(scores, k):
    """The reference's ranking rule: python stable sort, descending, slice."""
    hits = [{"id": i, "score": float(s)} for i, s in enumerate(scores)]
    hits.sort(key=lambda h: h["score"], reverse=True)
    return [h["id"] for h in hits[:k]]


def bf16_round(x):
    return bf16_bits_to_f32(f32_to_bf16_bits(np.asarray(x, dtype=np.float32)))


def gen_cosine_kat(rng):
    cases = []
    specials = [
        ([], [1.0]), ([1.0], []), ([], []), ([1.0, 2.0], [1.0]), ([0.0, 0.0], [1.0, 2.0]),
        ([1.0, 2.0], [0.0, 0.0]), ([0.1] * 8, [0.1] * 8), ([0.1] * 8, [0.3] * 8), ([1.0, 0.0], [0.0, 1.0]),
        ([1.0, 0.0], [-1.0, 0.0]), ([3.0, 4.0], [3.0, 4.0]), ([1e-200, 0.0], [1e-200, 0.0]),
    ]
    for a, b in specials:
        cases.append({"a": a, "b": b, "cos": ref_cosine(a, b)})
    for d in (3, 8, 48, 384, 768):
        for _ in range(4):
            a = rng.standard_normal(d).tolist()
            b = rng.standard_normal(d).tolist()
            cases.append({"a": a, "b": b, "cos": ref_cosine(a, b)})
    return cases


def gen_topk_fixture(rng, N, D, B, ks, dtype):
    X = rng.standard_normal((N, D)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    # adversarial structure: exact duplicates (ties), scaled duplicates (cos ties up to rounding),
    # zero rows (score 0.0), a zero query (all scores 0.0 -> ids 0..k-1)
    X[7] = X[3]
    X[N - 1] = X[3]
    X[N // 2] = X[3]
    X[11] = 0.0
    X[12] = 0.0
    X[20] = X[21] * 2.0
    Q[B - 1] = 0.0
    Q[B - 2] = X[3]
    if dtype == "bf16":
        X, Q = bf16_round(X), bf16_round(Q)
    elif dtype == "f16":
        X, Q = X.astype(np.float16).astype(np.float32), Q.astype(np.float16).astype(np.float32)
    S = np.zeros((B, N), dtype=np.float64)
    for b in range(B):
        qb = [float(v) for v in Q[b]]
        for n in range(N):
            S[b, n] = ref_cosine(qb, [float(v) for v in X[n]])
    out = {"X": X, "Q": Q, "S": S}
    for k in ks:
        out[f"ids_k{k}"] = np.array([ref_sorted_topk(S[b], k) for b in range(B)], dtype=np.int64)
    return out


def random_graph(rng, V, E, n_seeds):
    nodes = [{"id": "q1", "type": "question", "props": {"text": "what"}}]
    for i in range(V):
        nodes.append({"id": f"n{i}", "type": "sentence", "props": {"text": f"t{i}", "doc": f"d{i // 5}"}})
    edges = []
    src = rng.integers(0, V, size=E)
    dst = rng.integers(0, V, size=E)
    for s, t in zip(src, dst):
        edges.append({"source": f"n{s}", "target": f"n{t}", "type": "next_in_doc"})
    # edge types the reference BFS must ignore (graph_utils.py:49-53)
    for s, t in zip(rng.integers(0, V, size=E // 2), rng.integers(0, V, size=E // 2)):
        edges.append({"source": f"n{s}", "target": f"n{t}", "type": "semantic_sim"})
    seeds = sorted(set(int(s) for s in rng.integers(0, V, size=n_seeds)))
    for s in seeds:
        edges.append({"source": "q1", "target": f"n{s}", "type": "q_match"})
    edges.append({"source": "q2", "target": "n0", "type": "q_match"})  # wrong source: ignored (:52)
    return {"nodes": nodes, "edges": edges}, src.tolist(), dst.tolist(), seeds


def ref_expand(graph, window):
    nodes_by_id, nf, nb, node_texts, qmatch = build_index(graph)
    res = expand_qmatch_neighbors("what", nodes_by_id, nf, nb, node_texts, explicit_qmatch=qmatch, window=window)
    return {int(sid[1:]): (float(score), int(meta["distance"])) for sid, (score, meta) in res.items()}


def gen_bfs(rng):
    cases = []
    for V, E, ns in ((12, 10, 2), (60, 80, 4), (200, 150, 6), (300, 900, 20), (50, 0, 3)):
        graph, src, dst, seeds = random_graph(rng, V, E, ns)
        for window in (0, 1, 2, 3, 5):
            exp = ref_expand(graph, window)
            cases.append({"V": V, "src": src, "dst": dst, "seeds": seeds, "window": window,
                          "nodes": sorted(exp), "score": [exp[n][0] for n in sorted(exp)],
                          "dist": [exp[n][1] for n in sorted(exp)]})
    # a chain: the classic +/- window neighbourhood of a sentence
    V = 30
    graph = {"nodes": [{"id": f"n{i}", "type": "sentence", "props": {}} for i in range(V)],
             "edges": [{"source": f"n{i}", "target": f"n{i + 1}", "type": "next_in_doc"} for i in range(V - 1)]
             + [{"source": "q1", "target": "n10", "type": "q_match"}, {"source": "q1", "target": "n12", "type": "q_match"}]}
    for window in (1, 2, 4, 9):
        exp = ref_expand(graph, window)
        cases.append({"V": V, "src": list(range(V - 1)), "dst": list(range(1, V)), "seeds": [10, 12], "window": window,
                      "nodes": sorted(exp), "score": [exp[n][0] for n in sorted(exp)],
                      "dist": [exp[n][1] for n in sorted(exp)]})
    return cases


def gen_minmax(rng):
    cases = []
    for vals in ({}, {"a": 1.0}, {"a": 2.0, "b": 2.0}, {"a": 0.5, "b": 0.7, "c": 1.0},
                 {str(i): float(v) for i, v in enumerate(rng.standard_normal(17))}):
        cases.append({"in": vals, "out": ref_minmax(vals)})
    return cases


def gen_expand_rescore(rng):
    """Compose the reference's own pieces on an integer id space where channels join on id."""
    cases = []
    for V, E, ns, D, window, k in ((80, 120, 5, 32, 2, 10), (300, 900, 20, 64, 2, 10), (300, 400, 8, 64, 1, 5),
                                   (40, 30, 3, 32, 3, 50), (64, 100, 6, 32, 0, 10)):
        graph, src, dst, seeds = random_graph(rng, V, E, ns)
        X = bf16_round(rng.standard_normal((V, D)))
        X[5] = X[6]  # tie
        X[9] = 0.0   # zero-norm candidate
        q = bf16_round(rng.standard_normal(D))
        exp = ref_expand(graph, window)
        cand = sorted(exp)
        dense = {str(c): ref_cosine([float(v) for v in q], [float(v) for v in X[c]]) for c in cand}
        gsc = {str(c): exp[c][0] for c in cand}
        nd, ng = ref_minmax(dense), ref_minmax(gsc)
        a_d, a_g = 0.4, 0.2  # reference defaults (retrieval_backend.py:263-264); text channel absent -> 0
        fused = []
        for c in cand:  # id order, see SURVEY Appendix A.5
            ts, gs, ds = 0.0, ng.get(str(c), 0.0), nd.get(str(c), 0.0)
            fused.append({"id": c, "score": float(0.4 * ts + a_g * gs + a_d * ds), "g": gs, "d": ds})
        fused.sort(key=lambda h: h["score"], reverse=True)
        fused = fused[:k]
        cases.append({"V": V, "src": src, "dst": dst, "seeds": seeds, "window": window, "k": k, "D": D,
                      "alpha_dense": a_d, "alpha_graph": a_g,
                      "X_bf16": f32_to_bf16_bits(X).tolist(), "q_bf16": f32_to_bf16_bits(q).tolist(),
                      "cand": cand, "dist": [exp[c][1] for c in cand], "dense": [dense[str(c)] for c in cand],
                      "top_ids": [h["id"] for h in fused], "top_scores": [h["score"] for h in fused],
                      "top_g": [h["g"] for h in fused], "top_d": [h["d"] for h in fused]})
    return cases


def gen_mmr(rng):
    """mmr_diversify (similarity.py:33-62) on bf16-representable vectors: clustered items (so diversity matters), exact
    score ties, a None vector, a zero vector, top_k beyond the list, lambda 0 / 0.3 / 0.7 / 1."""
    from app.utils.similarity import mmr_diversify
    cases = []
    for n, D, top_k, lam in ((12, 16, 5, 0.7), (40, 32, 10, 0.3), (25, 24, 40, 0.7), (30, 16, 8, 0.0), (30, 16, 8, 1.0), (1, 8, 3, 0.5)):
        centers = rng.standard_normal((max(2, n // 6), D)).astype(np.float32)
        X = centers[rng.integers(0, len(centers), n)] + 0.15 * rng.standard_normal((n, D)).astype(np.float32)
        X = bf16_bits_to_f32(f32_to_bf16_bits(X))
        scores = np.round(rng.uniform(0.0, 1.0, n), 2).astype(np.float32)       # two decimals -> exact ties occur
        vec_none = int(rng.integers(0, n)) if n > 3 else -1
        vec_zero = int(rng.integers(0, n)) if n > 5 else -1
        if vec_zero >= 0:
            X[vec_zero] = 0.0
        ids = (rng.permutation(1000)[:n] + 7).tolist()
        items = [(ids[i], float(scores[i]), None if i == vec_none else X[i].astype(float).tolist()) for i in range(n)]
        picked = mmr_diversify(items, top_k=top_k, lambda_weight=lam)
        cases.append({"ids": ids, "scores": [float(s) for s in scores], "X_bf16": f32_to_bf16_bits(X).tolist(), "none": vec_none,
                      "top_k": top_k, "lambda": lam, "picked": [p[0] for p in picked]})
    return cases


def main():
    rng = np.random.default_rng(20261018)
    with open(os.path.join(HERE, "cosine_kat.json"), "w") as f:
        json.dump(gen_cosine_kat(rng), f)
    np.savez_compressed(os.path.join(HERE, "topk_bf16_n600_d64.npz"), **gen_topk_fixture(rng, 600, 64, 6, (1, 10, 100), "bf16"))
    np.savez_compressed(os.path.join(HERE, "topk_f32_n400_d48.npz"), **gen_topk_fixture(rng, 400, 48, 5, (10, 100), "f32"))
    np.savez_compressed(os.path.join(HERE, "topk_f16_n300_d384.npz"), **gen_topk_fixture(rng, 300, 384, 4, (10, 300), "f16"))
    with open(os.path.join(HERE, "bfs_cases.json"), "w") as f:
        json.dump(gen_bfs(rng), f)
    with open(os.path.join(HERE, "minmax_cases.json"), "w") as f:
        json.dump(gen_minmax(rng), f)
    with open(os.path.join(HERE, "expand_rescore_cases.json"), "w") as f:
        json.dump(gen_expand_rescore(rng), f)
    with open(os.path.join(HERE, "mmr_cases.json"), "w") as f:
        json.dump(gen_mmr(np.random.default_rng(20261019)), f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
