"""CPU oracle for the retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (numpy + a few pure-Python loops) of the reference's
algorithm for the path named in BASELINE.json:north_star.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package (``a_modular_rag_framework_b200``) never does.

Parity pinning: the reference has NO tests / golden vectors of its own, so the
oracle is pinned against outputs of the reference's own functions, imported in the build
container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.json|npz``
(``tests/test_oracle_golden.py`` replays them).

Reference lines each function follows (paths relative to the reference repo root):

* ``cosine_pair``            app/modules/retrieval/retrieval_backend.py:192-197  (DenseReranker._cosine)
* ``query_limbs_int8`` / ``int8_limb_scores``  no reference counterpart: restatement of OUR int8 arithmetic, so that its
                             error against ``score_matrix`` can be bounded on the CPU (tests/test_int8_limbs.py)
* ``score_matrix`` / ``topk``  same metric, vectorised; top-k = stable ``sort(reverse=True)[:k]`` of
                             retrieval_backend.py:371-372, text_index.py:96, retrieval_adapter.py:129-131
                             (stable descending sort over rows enumerated in id order == ties to lower id)
* ``merge_topk``             same sort rule applied to the union of per-shard top-k lists
* ``decay`` / ``bfs_expand``  app/modules/retrieval/graph_utils.py:87-94, 109-127
* ``minmax_norm``            retrieval_backend.py:296-301
* ``fuse_and_rank``          retrieval_backend.py:357-372 (weighted sum, stable sort, slice)
* ``expand_rescore``         composition a4+a1+a6+a3 of SURVEY.md section 8(a)
* ``mmr_diversify``          app/utils/similarity.py:33-62 (+ cosine :11-16)
"""
from __future__ import annotations

import math
from collections import deque
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------- dtype helpers


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounding = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + rounding) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def quantize_int8_rows(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row absmax int8 quantisation: x ~= q * scale[:, None]."""
    x = np.asarray(x, dtype=np.float32)
    amax = np.abs(x).max(axis=1)
    scale = np.where(amax > 0, amax / 127.0, 1.0).astype(np.float32)
    q = np.clip(np.rint(x / scale[:, None]), -127, 127).astype(np.int8)
    return q, scale


# ----------------------------------------------------------------------------- a1: the metric


def query_limbs_int8(q: np.ndarray) -> Tuple[np.ndarray, float]:
    """The int8 paths' image of one fp32 query (csrc/rows_i8.cu quantize_queries_i8_kernel, csrc/scan_bulk.cu):
    23-bit fixed point relative to the largest component, split into three balanced signed-byte limbs.
    -> (limbs int8 [3, D] with q_int = l2*65536 + l1*256 + l0, scale) such that q ~= scale * q_int and
    |q - scale*q_int| <= 2^-22 * max|q| (half a grid step from the rounding plus the fp32 rounding of the scaling).  No counterpart in the reference (it has no quantised path); the parity bar
    for int8 corpora is still the fp64 cosine of retrieval_backend.py:192-197 on the dequantised rows."""
    q = np.asarray(q, dtype=np.float32)
    qmax = np.float32(np.max(np.abs(q))) if q.size else np.float32(0)
    if not qmax > 0:
        return np.zeros((3, q.shape[0]), dtype=np.int8), 0.0
    inv = np.float32(4194304.0) / qmax
    qi = np.rint(q * inv).astype(np.int64)              # round-half-even like __float2int_rn, |qi| <= 2^22
    l0 = ((qi + 128) & 255) - 128
    q1 = (qi - l0) >> 8
    l1 = ((q1 + 128) & 255) - 128
    l2 = (q1 - l1) >> 8
    limbs = np.stack([l0, l1, l2])
    assert limbs.min() >= -128 and limbs.max() <= 127
    return limbs.astype(np.int8), float(qmax * np.float32(1.0 / 4194304.0))


def int8_limb_scores(q: np.ndarray, X8: np.ndarray, row_mult: Optional[np.ndarray] = None) -> np.ndarray:
    """Scores exactly as the int8 tensor path forms them: exact integer limb sums, fp32 recombination
    a2*65536 + a1*256 + a0, times the query scale, times the row multiplier."""
    limbs, scale = query_limbs_int8(q)
    A = X8.astype(np.int64) @ limbs.astype(np.int64).T          # [N, 3] exact
    assert np.abs(A).max(initial=0) < 2 ** 31
    f = np.float32(65536.0) * A[:, 2].astype(np.float32)
    f = (f + (np.float32(256.0) * A[:, 1].astype(np.float32) + A[:, 0].astype(np.float32)).astype(np.float32)).astype(np.float32)
    s = (f * np.float32(scale)).astype(np.float32)
    if row_mult is not None:
        s = (s * row_mult.astype(np.float32)).astype(np.float32)
    return s


def cosine_pair(a: Sequence[float], b: Sequence[float]) -> float:
    """Literal restatement of DenseReranker._cosine (retrieval_backend.py:192-197)."""
    if len(a) == 0 or len(b) == 0 or len(a) != len(b):
        return 0.0
    # NB: builtin sum() over floats is Neumaier-compensated on CPython >= 3.12 (the interpreter on
    # both boxes); the fixtures were recorded under it, so use the same primitive as the reference.
    dot = sum(x * y for x, y in zip(a, b))
    na = sum(x * x for x in a)
    nb = sum(y * y for y in b)
    na = math.sqrt(na)
    nb = math.sqrt(nb)
    return float(dot / (na * nb)) if na and nb else 0.0


def score_matrix(Q: np.ndarray, X: np.ndarray, metric: str = "cos", dtype=np.float64) -> np.ndarray:
    """S[b, n] = cos(Q[b], X[n]) (or the inner product), zero-norm -> 0.0 like :197."""
    Qd = np.asarray(Q, dtype=dtype)
    Xd = np.asarray(X, dtype=dtype)
    S = Qd @ Xd.T
    if metric == "ip":
        return S
    if metric != "cos":
        raise ValueError(f"unknown metric {metric!r}")
    qn = np.sqrt((Qd * Qd).sum(axis=1))
    xn = np.sqrt((Xd * Xd).sum(axis=1))
    denom = qn[:, None] * xn[None, :]
    out = np.zeros_like(S)
    np.divide(S, denom, out=out, where=denom != 0)
    return out


# ----------------------------------------------------------------------------- a3: top-k


def topk_from_scores(S: np.ndarray, k: int, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Stable descending sort + slice over rows enumerated in id order (ties -> lower id)."""
    B, N = S.shape
    kk = min(k, N)
    ids = np.empty((B, kk), dtype=np.int64)
    sc = np.empty((B, kk), dtype=S.dtype)
    for b in range(B):
        order = np.argsort(-S[b], kind="stable")[:kk]
        ids[b] = order + id_base
        sc[b] = S[b, order]
    return sc, ids


def topk(Q: np.ndarray, X: np.ndarray, k: int, metric: str = "cos", block_rows: int = 1 << 18,
         dtype=np.float64, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k of every query against every row, blockwise so large N fits in RAM.

    Returns (scores[B, min(k,N)], ids[B, min(k,N)] int64) ordered by (score desc, id asc).
    """
    Q = np.asarray(Q)
    B = Q.shape[0]
    N = X.shape[0]
    best_s = np.empty((B, 0), dtype=dtype)
    best_i = np.empty((B, 0), dtype=np.int64)
    for r0 in range(0, N, block_rows):
        r1 = min(N, r0 + block_rows)
        S = score_matrix(Q, X[r0:r1], metric, dtype)
        kk = min(k, r1 - r0)
        if kk < r1 - r0:
            # keep everything >= the kk-th value (ties included) so the exact merge can decide
            part = np.partition(S, (r1 - r0) - kk, axis=1)[:, (r1 - r0) - kk]
            cand_s, cand_i = [], []
            for b in range(B):
                sel = np.nonzero(S[b] >= part[b])[0]
                cand_s.append(S[b, sel])
                cand_i.append(sel + r0)
        else:
            cand_s = [S[b] for b in range(B)]
            cand_i = [np.arange(r0, r1, dtype=np.int64) for _ in range(B)]
        new_s, new_i = [], []
        width = 0
        for b in range(B):
            s = np.concatenate([best_s[b], cand_s[b]])
            i = np.concatenate([best_i[b], cand_i[b].astype(np.int64)])
            order = np.lexsort((i, -s))[:k]
            new_s.append(s[order])
            new_i.append(i[order])
            width = max(width, len(order))
        best_s = np.stack([np.pad(s, (0, width - len(s)), constant_values=-np.inf) for s in new_s])
        best_i = np.stack([np.pad(i, (0, width - len(i)), constant_values=-1) for i in new_i])
    return best_s, best_i + id_base


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard top-k lists: scores/ids [B, G, kk] -> [B, k] by (score desc, id asc).

    Entries with id < 0 are padding and never win.
    """
    B = scores.shape[0]
    s = scores.reshape(B, -1)
    i = ids.reshape(B, -1).astype(np.int64)
    out_s = np.full((B, k), -np.inf, dtype=s.dtype)
    out_i = np.full((B, k), -1, dtype=np.int64)
    for b in range(B):
        valid = i[b] >= 0
        sb, ib = s[b][valid], i[b][valid]
        order = np.lexsort((ib, -sb))[:k]
        out_s[b, : len(order)] = sb[order]
        out_i[b, : len(order)] = ib[order]
    return out_s, out_i


# ----------------------------------------------------------------------------- a4: graph expansion


def decay(d: int) -> float:
    """graph_utils.py:87-94."""
    if d == 0:
        return 1.0
    if d == 1:
        return 0.7
    if d == 2:
        return 0.5
    return max(0.5 - 0.1 * (d - 2), 0.1)


def csr_from_edges(num_nodes: int, src: np.ndarray, dst: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Undirected adjacency as the reference walks it (graph_utils.py:50-51,123):
    neighbours(s) = next_forward[s] + next_backward[s].  Parallel edges are kept (the BFS's
    ``visited`` set makes them harmless); neighbour lists are sorted by id."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    a = np.concatenate([src, dst])
    b = np.concatenate([dst, src])
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(row_ptr, a + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return row_ptr, b.astype(np.int32)


def bfs_expand(row_ptr: np.ndarray, col_idx: np.ndarray, seeds: Iterable[int], window: int) -> Dict[int, int]:
    """Multi-source BFS, depth <= window, global visited set => min hop distance
    (graph_utils.py:109-127).  Returns {node: distance}.  Negative seeds are padding."""
    window = max(0, int(window))
    seeds = [int(s) for s in seeds if int(s) >= 0]
    dist: Dict[int, int] = {}
    queue = deque()
    for s in dict.fromkeys(seeds):
        dist[s] = 0
        queue.append(s)
    while queue:
        u = queue.popleft()
        d = dist[u]
        if d >= window:
            continue
        for v in col_idx[row_ptr[u]: row_ptr[u + 1]]:
            v = int(v)
            if v not in dist:
                dist[v] = d + 1
                queue.append(v)
    return dist


# ----------------------------------------------------------------------------- a6: fusion


def minmax_norm(values: Dict[int, float]) -> Dict[int, float]:
    """retrieval_backend.py:296-301."""
    if not values:
        return {}
    vs = list(values.values())
    vmin, vmax = min(vs), max(vs)
    if vmax <= vmin:
        return {k: 0.0 for k in values}
    return {k: (v - vmin) / (vmax - vmin) for k, v in values.items()}


def fuse_and_rank(dense: Dict[int, float], graph: Dict[int, float], alpha_dense: float, alpha_graph: float,
                  k: int, text: Optional[Dict[int, float]] = None, alpha_text: float = 0.0
                  ) -> List[Tuple[int, float, float, float, float]]:
    """retrieval_backend.py:353-372 on an integer id space where the channels join on id.

    Returns [(id, fused, text_norm, graph_norm, dense_norm)] sorted by (fused desc, id asc), first k.
    Candidates are enumerated in id order, which is what makes the reference's stable sort
    deterministic.
    """
    nt = minmax_norm(text or {})
    ng = minmax_norm(graph)
    nd = minmax_norm(dense)
    ids = sorted(set(nt) | set(ng) | set(nd))
    fused = []
    for nid in ids:
        ts, gs, ds = nt.get(nid, 0.0), ng.get(nid, 0.0), nd.get(nid, 0.0)
        fused.append((nid, alpha_text * ts + alpha_graph * gs + alpha_dense * ds, ts, gs, ds))
    fused.sort(key=lambda h: h[1], reverse=True)
    return fused[:k]


def expand_rescore(q: np.ndarray, X: np.ndarray, row_ptr: np.ndarray, col_idx: np.ndarray,
                   seeds: Iterable[int], window: int, alpha_dense: float, alpha_graph: float, k: int,
                   metric: str = "cos", node_rows: Optional[np.ndarray] = None
                   ) -> Tuple[List[Tuple[int, float, float, float, float]], Dict[int, int]]:
    """Graph-enhanced second stage for ONE query: BFS candidates, cosine re-score, fuse, top-k.

    node_rows (optional): the graph's node ids are local and node_rows[v] is node v's corpus row (the shape of the
    reference's per-graph_id graphs, graph_utils.py:13-55, whose node ids are strings); seeds are then local ids and the
    returned ids / distance keys are corpus rows.  A node without a corpus row (or beyond X) scores 0.0."""
    dist = bfs_expand(row_ptr, col_idx, seeds, window)
    if node_rows is not None:
        dist = {int(node_rows[v]): d for v, d in dist.items()}
    cand = np.array(sorted(dist), dtype=np.int64)
    if len(cand) == 0:
        return [], dist
    X = np.asarray(X)
    Xc = np.zeros((len(cand), X.shape[1]), dtype=np.float64)
    inside = (cand >= 0) & (cand < X.shape[0])
    Xc[inside] = X[cand[inside]]
    S = score_matrix(np.asarray(q)[None, :], Xc, metric)[0]
    dense = {int(c): float(s) for c, s in zip(cand, S)}
    graph = {int(c): decay(dist[int(c)]) for c in cand}
    return fuse_and_rank(dense, graph, alpha_dense, alpha_graph, k), dist


def mmr_diversify(items: Sequence[Tuple[int, float, Optional[Sequence[float]]]], top_k: int = 20, lambda_weight: float = 0.7
                  ) -> List[Tuple[int, float, Optional[Sequence[float]]]]:
    """Restatement of mmr_diversify (app/utils/similarity.py:33-62): greedy MMR; the first pick maximises the raw score, later
    picks lambda*score - (1-lambda)*sim_max with sim_max starting at 0.0 (negative cosines never count) and cosine as in
    similarity.py:11-16 (0.0 for empty / zero vectors); `val > best_val` is strict, so ties go to the earlier item."""
    def cos(u, v):
        if u is None or v is None or len(u) == 0 or len(v) == 0:
            return 0.0
        du = math.sqrt(sum(x * x for x in u))
        dv = math.sqrt(sum(x * x for x in v))
        if du == 0 or dv == 0:
            return 0.0
        return sum(x * y for x, y in zip(u, v)) / (du * dv)
    selected: List[Tuple[int, float, Optional[Sequence[float]]]] = []
    candidates = list(items)
    while candidates and len(selected) < top_k:
        best, best_val = None, -1e9
        for cid, cscore, cvec in candidates:
            if not selected:
                val = cscore
            else:
                sim_max = 0.0
                for _, _, svec in selected:
                    sim_max = max(sim_max, cos(cvec, svec) if (cvec is not None and svec is not None) else 0.0)
                val = lambda_weight * cscore - (1 - lambda_weight) * sim_max
            if val > best_val:
                best_val, best = val, (cid, cscore, cvec)
        selected.append(best)
        candidates = [x for x in candidates if x[0] != best[0]]
    return selected


# ----------------------------------------------------------------------------- comparison helper


def check_topk(got_scores: np.ndarray, got_ids: np.ndarray, ref_full_scores: np.ndarray, k: int,
               rel_tol: float, id_base: int = 0) -> None:
    """Tie-aware parity check of one query's result against the oracle's FULL score row.

    * ids are unique, valid, and exactly min(k, N) of them are returned;
    * every returned id's oracle score is within tolerance of the returned score;
    * the returned set is the oracle top-k up to near-ties at the k-th boundary
      (|s - s_k| <= tol); among EXACT oracle ties the lower id must win;
    * the returned list is ordered by score descending.
    """
    N = ref_full_scores.shape[0]
    kk = min(k, N)
    ids = np.asarray(got_ids[:kk], dtype=np.int64) - id_base
    sc = np.asarray(got_scores[:kk], dtype=np.float64)
    assert ids.min() >= 0 and ids.max() < N, "id out of range"
    assert len(np.unique(ids)) == kk, "duplicate ids returned"
    ref = ref_full_scores.astype(np.float64)
    scale = max(1e-30, float(np.max(np.abs(ref))))
    tol = rel_tol * scale
    assert np.all(np.abs(ref[ids] - sc) <= tol + rel_tol * np.abs(ref[ids])), "score mismatch"
    assert np.all(sc[:-1] >= sc[1:] - 2 * tol), "not sorted by score"
    order = np.lexsort((np.arange(N), -ref))
    s_k = ref[order[kk - 1]]
    must = set(np.nonzero(ref > s_k + 2 * tol)[0].tolist())
    may = set(np.nonzero(ref >= s_k - 2 * tol)[0].tolist())
    got = set(ids.tolist())
    assert must <= got, f"missing {sorted(must - got)[:5]}"
    assert got <= may, f"spurious {sorted(got - may)[:5]}"


# ----------------------------------------------------------------------------- large-N verification (bench.py's recall leg)


def topk_blocks_exact(Q: np.ndarray, blocks: Iterable[Tuple[int, np.ndarray]], k: int, metric: str = "cos",
                      extra: int = 64, margin: float = 2e-4) -> Tuple[np.ndarray, np.ndarray]:
    """Exact (fp64) top-(k+extra) of a few queries over a corpus that arrives in row blocks ``(first_id, X_block f32)``,
    at sizes where a full fp64 score matrix per block would take minutes (10M x 768).

    Same metric and ranking rule as ``topk`` (retrieval_backend.py:192-197, 371-372).  Per block the float32 BLAS scores
    only PRE-FILTER: every row whose f32 score is within ``margin`` (absolute; >= 4x the worst-case f32 rounding of a
    768-term dot product of unit-scale vectors) of the block's (k+extra)-th f32 score is kept, and the kept rows are
    re-scored in fp64 exactly like ``score_matrix``; the running result is the exact (score desc, id asc) order of those
    fp64 scores.  A row of the true top-(k+extra) cannot be dropped: its f32 score is within margin/4 of its fp64 score.
    Returns (scores f64 [B, k+extra], ids i64 [B, k+extra]) padded with (-inf, -1)."""
    Q64 = np.asarray(Q, dtype=np.float64)
    Q32 = np.asarray(Q, dtype=np.float32)
    B = Q64.shape[0]
    want = k + extra
    qn = np.sqrt((Q64 * Q64).sum(axis=1))
    best_s = [np.empty(0, dtype=np.float64) for _ in range(B)]
    best_i = [np.empty(0, dtype=np.int64) for _ in range(B)]
    for first_id, Xb in blocks:
        Xb = np.asarray(Xb, dtype=np.float32)
        n = Xb.shape[0]
        if n == 0:
            continue
        S = Q32 @ Xb.T                                                  # [B, n] prefilter only
        if metric == "cos":
            xn32 = np.sqrt(np.einsum("nd,nd->n", Xb, Xb))
            qn32 = np.sqrt((Q32 * Q32).sum(axis=1))
            den = qn32[:, None] * xn32[None, :]
            np.divide(S, den, out=S, where=den != 0)
            S[den == 0] = 0.0
        kk = min(want, n)
        for b in range(B):
            if kk < n:
                cut = np.partition(S[b], n - kk)[n - kk] - margin
                sel = np.nonzero(S[b] >= cut)[0]
            else:
                sel = np.arange(n)
            Xs = Xb[sel].astype(np.float64)
            s = Xs @ Q64[b]
            if metric == "cos":
                xn = np.sqrt((Xs * Xs).sum(axis=1))
                den = xn * qn[b]
                s = np.divide(s, den, out=np.zeros_like(s), where=den != 0)
            cs = np.concatenate([best_s[b], s])
            ci = np.concatenate([best_i[b], sel.astype(np.int64) + int(first_id)])
            order = np.lexsort((ci, -cs))[:want]
            best_s[b], best_i[b] = cs[order], ci[order]
    out_s = np.full((B, want), -np.inf, dtype=np.float64)
    out_i = np.full((B, want), -1, dtype=np.int64)
    for b in range(B):
        out_s[b, : len(best_s[b])] = best_s[b]
        out_i[b, : len(best_i[b])] = best_i[b]
    return out_s, out_i


def check_topk_against_list(got_scores: np.ndarray, got_ids: np.ndarray, ref_scores: np.ndarray, ref_ids: np.ndarray,
                            k: int, rel_tol: float) -> Dict[str, float]:
    """``check_topk`` for ONE query when only the oracle's ordered top-(k+extra) list is available (large N).

    Tie-aware rule of the north star: with s_k the oracle's k-th score and tol = rel_tol * max|score|, every oracle
    entry above s_k + 2 tol must be returned, every returned id must be an oracle entry >= s_k - 2 tol, returned scores
    must agree with the oracle's within rel_tol (relative), and bit-equal returned scores must list the lower id first
    when the oracle ties them exactly.  Needs the list to reach below s_k - 2 tol (``deep_enough``); raises otherwise.
    Returns {"recall", "recall_strict", "max_rel_score_err", "ok"}."""
    valid = ref_ids >= 0
    rs, ri = ref_scores[valid].astype(np.float64), ref_ids[valid].astype(np.int64)
    kk = min(k, len(ri))
    gs, gi = np.asarray(got_scores[:kk], dtype=np.float64), np.asarray(got_ids[:kk], dtype=np.int64)
    scale = max(1e-30, float(np.max(np.abs(rs)))) if len(rs) else 1.0
    tol = rel_tol * scale
    s_k = rs[kk - 1]
    if len(rs) > kk and not rs[-1] < s_k - 2 * tol:
        raise ValueError("oracle list too shallow for the tie-aware check: raise `extra`")
    ref_of = {int(i): float(s) for i, s in zip(ri, rs)}
    must = {int(i) for i, s in zip(ri, rs) if s > s_k + 2 * tol}
    may = {int(i) for i, s in zip(ri, rs) if s >= s_k - 2 * tol}
    got = [int(i) for i in gi]
    gset = set(got)
    ok = len(gset) == kk and must <= gset and gset <= may
    hit = len(gset & may)
    errs = [abs(ref_of[i] - s) / scale for i, s in zip(got, gs) if i in ref_of]
    max_err = max(errs) if errs else float("inf")
    ok = ok and max_err <= rel_tol and bool(np.all(gs[:-1] >= gs[1:] - 2 * tol))
    for j in np.nonzero(gs[1:] == gs[:-1])[0]:
        a, c = got[j], got[j + 1]
        if a in ref_of and c in ref_of and ref_of[a] == ref_of[c] and not a < c:
            ok = False
    return {"recall": hit / kk if kk else 1.0, "recall_strict": len(gset & set(ri[:kk].tolist())) / kk if kk else 1.0,
            "max_rel_score_err": max_err, "ok": bool(ok)}
