"""The oracle restatement against the fixtures recorded from the reference's own functions
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import retrieval_oracle as O


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_cosine_pair_bit_exact(golden_dir):
    for c in _load(golden_dir, "cosine_kat.json"):
        assert O.cosine_pair(c["a"], c["b"]) == c["cos"]


def test_score_matrix_matches_reference_cosine(golden_dir):
    for c in _load(golden_dir, "cosine_kat.json"):
        if len(c["a"]) != len(c["b"]) or not c["a"]:
            continue
        s = O.score_matrix(np.array([c["a"]]), np.array([c["b"]]))[0, 0]
        assert abs(s - c["cos"]) <= 1e-12 * max(1.0, abs(c["cos"]))


@pytest.mark.parametrize("name", ["topk_bf16_n600_d64.npz", "topk_f32_n400_d48.npz", "topk_f16_n300_d384.npz"])
def test_topk_matches_reference_sort(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    X, Q, S = z["X"], z["Q"], z["S"]
    S_or = O.score_matrix(Q, X)
    assert np.max(np.abs(S_or - S)) <= 1e-12
    for key in z.files:
        if not key.startswith("ids_k"):
            continue
        k = int(key[5:])
        # ranking rule alone, on the reference's own scores: must be identical ids in identical order
        _, ids = O.topk_from_scores(S, k)
        assert np.array_equal(ids, z[key])
        # blocked top-k (different block sizes) on the oracle's scores: tie-aware check vs reference scores
        for block in (64, 1 << 18):
            sc, ids2 = O.topk(Q, X, k, block_rows=block)
            for b in range(Q.shape[0]):
                O.check_topk(sc[b], ids2[b], S[b], k, rel_tol=1e-11)


def test_topk_exact_ties_prefer_lower_id():
    X = np.ones((50, 8), dtype=np.float32)
    Q = np.ones((2, 8), dtype=np.float32)
    sc, ids = O.topk(Q, X, 10, block_rows=7)
    assert np.array_equal(ids, np.tile(np.arange(10), (2, 1)))
    Z = np.zeros((2, 8), dtype=np.float32)
    sc, ids = O.topk(Z, X, 5)
    assert np.array_equal(ids, np.tile(np.arange(5), (2, 1))) and np.all(sc == 0.0)


def test_merge_topk_equals_unsharded():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1000, 16)).astype(np.float32)
    X[500] = X[2]
    Q = rng.standard_normal((4, 16)).astype(np.float32)
    k = 20
    full_s, full_i = O.topk(Q, X, k)
    parts_s, parts_i = [], []
    for g in range(4):
        s, i = O.topk(Q, X[g * 250:(g + 1) * 250], k, id_base=g * 250)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = O.merge_topk(np.stack(parts_s, 1), np.stack(parts_i, 1), k)
    assert np.array_equal(mi, full_i)
    assert np.allclose(ms, full_s, rtol=0, atol=0)


def test_bfs_matches_reference(golden_dir):
    for c in _load(golden_dir, "bfs_cases.json"):
        row_ptr, col_idx = O.csr_from_edges(c["V"], np.array(c["src"], dtype=np.int64), np.array(c["dst"], dtype=np.int64))
        dist = O.bfs_expand(row_ptr, col_idx, c["seeds"], c["window"])
        assert sorted(dist) == c["nodes"]
        assert [dist[n] for n in c["nodes"]] == c["dist"]
        assert [O.decay(dist[n]) for n in c["nodes"]] == c["score"]


def test_minmax_matches_reference(golden_dir):
    for c in _load(golden_dir, "minmax_cases.json"):
        assert O.minmax_norm(c["in"]) == c["out"]


def test_expand_rescore_matches_reference_composition(golden_dir):
    for c in _load(golden_dir, "expand_rescore_cases.json"):
        X = O.bf16_bits_to_f32(np.array(c["X_bf16"], dtype=np.uint16))
        q = O.bf16_bits_to_f32(np.array(c["q_bf16"], dtype=np.uint16))
        row_ptr, col_idx = O.csr_from_edges(c["V"], np.array(c["src"], dtype=np.int64), np.array(c["dst"], dtype=np.int64))
        top, dist = O.expand_rescore(q, X, row_ptr, col_idx, c["seeds"], c["window"], c["alpha_dense"], c["alpha_graph"], c["k"])
        assert sorted(dist) == c["cand"]
        assert [dist[n] for n in c["cand"]] == c["dist"]
        assert [t[0] for t in top] == c["top_ids"]
        assert np.allclose([t[1] for t in top], c["top_scores"], rtol=1e-12, atol=1e-13)
        assert np.allclose([t[3] for t in top], c["top_g"], rtol=1e-12, atol=1e-13)
        assert np.allclose([t[4] for t in top], c["top_d"], rtol=1e-10, atol=1e-12)


def test_bf16_roundtrip_helpers():
    x = np.array([1.0, -2.5, 3.1415927, 1e-3, 65504.0, 0.0], dtype=np.float32)
    b = O.f32_to_bf16_bits(x)
    y = O.bf16_bits_to_f32(b)
    assert np.all(np.abs(y - x) <= np.abs(x) * 2 ** -8)
    assert np.array_equal(O.f32_to_bf16_bits(y), b)


def test_blockwise_prefiltered_topk_equals_exact_topk():
    """bench.py's recall leg uses topk_blocks_exact (f32 prefilter + fp64 re-score per row block): it must reproduce
    the plain fp64 oracle, ties (duplicated rows across blocks) included."""
    rng = np.random.default_rng(5)
    N, D, B, k = 9000, 96, 5, 40
    X = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((N, D)).astype(np.float32)))
    X[7000] = X[11]
    X[N - 1] = X[11]
    X[123] = 0.0
    Q = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((B, D)).astype(np.float32)))
    Q[0] = X[11]
    for metric in ("cos", "ip"):
        ref_s, ref_i = O.topk(Q, X, k + 8, metric=metric)
        blocks = ((r0, X[r0:r0 + 2048]) for r0 in range(0, N, 2048))
        s, i = O.topk_blocks_exact(Q, blocks, k, metric=metric, extra=8)
        assert np.array_equal(i, ref_i)
        assert np.allclose(s, ref_s, rtol=0, atol=1e-12)
        for b in range(B):
            r = O.check_topk_against_list(ref_s[b, :k].astype(np.float32), ref_i[b, :k], s[b], i[b], k, 1e-3)
            assert r["ok"] and r["recall"] == 1.0 and r["recall_strict"] == 1.0 and r["max_rel_score_err"] < 1e-6
        # a wrong id (outside the near-tie band) and a wrong score must both be caught
        bad_i = ref_i[1, :k].copy()
        bad_i[3] = int(i[1, -1]) if i[1, -1] not in bad_i else N + 5
        if s[1, -1] < s[1, k - 1] - 1e-2:
            assert not O.check_topk_against_list(ref_s[1, :k], bad_i, s[1], i[1], k, 1e-3)["ok"]
        bad_s = ref_s[1, :k].copy()
        bad_s[0] *= 1.01
        assert not O.check_topk_against_list(bad_s, ref_i[1, :k], s[1], i[1], k, 1e-3)["ok"]


def test_blockwise_topk_shard_merge_equals_global():
    """Per-rank lists (id_base = shard start) merged with merge_topk equal the single-process list: what bench.py's recall
    leg does at N > 1."""
    rng = np.random.default_rng(6)
    N, D, B, k = 5000, 32, 3, 20
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    full_s, full_i = O.topk_blocks_exact(Q, [(0, X)], k, extra=4)
    parts = [O.topk_blocks_exact(Q, [(lo, X[lo:lo + 1700])], k, extra=4) for lo in range(0, N, 1700)]
    ms, mi = O.merge_topk(np.stack([p[0] for p in parts], axis=1), np.stack([p[1] for p in parts], axis=1), k + 4)
    assert np.array_equal(mi, full_i) and np.array_equal(ms, full_s)


def test_mmr_restatement_matches_reference(golden_dir):
    """oracle.mmr_diversify against picks recorded from the reference's mmr_diversify (app/utils/similarity.py:33-62)."""
    for c in _load(golden_dir, "mmr_cases.json"):
        X = O.bf16_bits_to_f32(np.array(c["X_bf16"], dtype=np.uint16))
        items = [(i, s, None if j == c["none"] else X[j].astype(float).tolist()) for j, (i, s) in enumerate(zip(c["ids"], c["scores"]))]
        got = O.mmr_diversify(items, top_k=c["top_k"], lambda_weight=c["lambda"])
        assert [g[0] for g in got] == c["picked"]
