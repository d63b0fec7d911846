"""Shared parity check of the graph stage (BFS + re-score + fusion + top-k) against the oracle.

Bar (BASELINE.json north_star): hop distances and candidate counts are integers -> exact; the raw cosine agrees within
1e-3 RELATIVE (bf16/fp16 inputs; 1e-5 for fp32) of the oracle's fp64 cosine; the fused score agrees within the bound that
cosine tolerance implies through the min-max map (alpha_dense * 2 * tol_cos / (max - min)); ids equal the oracle's ranking
position by position, except inside a NEAR-TIE: position i may hold any candidate whose oracle fused score is within that
fused tolerance of the oracle's i-th score (and exact oracle ties must list the lower id first)."""
import numpy as np

from oracle import retrieval_oracle as O

COS_REL = {"bf16": 1e-3, "f16": 1e-3, "f32": 1e-5}


def oracle_graph_query(q, X, row_ptr, col_idx, seeds, window, a_d, a_g):
    """-> (full ranking [(id, fused, t, g, d)], {node: hop}, {node: fp64 cosine})."""
    dist = O.bfs_expand(row_ptr, col_idx, seeds, window)
    cand = np.array(sorted(dist), dtype=np.int64)
    if len(cand) == 0:
        return [], dist, {}
    Xc = np.zeros((len(cand), X.shape[1]), dtype=np.float64)
    inside = cand < X.shape[0]                      # graph nodes without a corpus row score 0.0
    Xc[inside] = X[cand[inside]]
    cos = O.score_matrix(np.asarray(q, dtype=np.float64)[None, :], Xc, "cos")[0]
    dense = {int(c): float(s) for c, s in zip(cand, cos)}
    graph = {int(c): O.decay(dist[int(c)]) for c in cand}
    return O.fuse_and_rank(dense, graph, a_d, a_g, len(cand)), dist, dense


def check_graph_query(out, b, q, X, row_ptr, col_idx, seeds, window, a_d, a_g, k, kind="bf16"):
    full, dist, dense = oracle_graph_query(q, X, row_ptr, col_idx, seeds, window, a_d, a_g)
    ids = out["id"][b].cpu().numpy()
    sc = out["score"][b].cpu().numpy().astype(np.float64)
    hop = out["hop"][b].cpu().numpy()
    cosv = out["cos"][b].cpu().numpy().astype(np.float64)
    n = min(k, len(full))
    assert int(out["ncand"][b]) == len(dist)
    assert np.all(ids[n:] == -1) and np.all(np.isneginf(sc[n:]))
    if n == 0:
        return
    vals = np.array(list(dense.values()))
    scale = max(1e-30, float(np.max(np.abs(vals))))
    tol_cos = COS_REL[kind] * scale
    rng_ = float(vals.max() - vals.min())
    tol_f = (abs(a_d) * 2.0 * tol_cos / rng_ if rng_ > 0 else 0.0) + 1e-6
    ref_sc = np.array([t[1] for t in full])
    ref_of = {t[0]: t[1] for t in full}
    assert len(set(ids[:n].tolist())) == n, "duplicate ids"
    for i in range(n):
        nid = int(ids[i])
        assert nid in dist, f"id {nid} is not a BFS candidate"
        assert hop[i] == dist[nid]
        assert abs(cosv[i] - dense[nid]) <= tol_cos, (i, nid, cosv[i], dense[nid])
        assert abs(sc[i] - ref_of[nid]) <= tol_f, (i, nid, sc[i], ref_of[nid])
        if nid != full[i][0]:
            assert abs(ref_of[nid] - ref_sc[i]) <= tol_f, f"position {i}: got {nid}, oracle {full[i][0]} (not a near-tie)"
    assert np.all(sc[:n - 1] >= sc[1:n]), "not sorted by fused score"
    for j in np.nonzero(sc[1:n] == sc[:n - 1])[0]:          # bit-equal fused scores: exact oracle ties -> lower id first
        a, c = int(ids[j]), int(ids[j + 1])
        if ref_of[a] == ref_of[c]:
            assert a < c
    # nothing outside the near-tie band of the k-th score may be missing
    if len(full) > n:
        must = {t[0] for t in full if t[1] > ref_sc[n - 1] + tol_f}
        assert must <= set(ids[:n].tolist())
