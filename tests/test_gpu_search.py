"""Parity of the CUDA search paths (through the C ABI) against the oracle and the golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 1e-3, torch.float16: 1e-3, torch.float32: 1e-5, torch.int8: 1e-5}


def _dev():
    return torch.device("cuda:0")


def _to_dev(x_f32: np.ndarray, dtype):
    return torch.from_numpy(np.ascontiguousarray(x_f32, dtype=np.float32)).to(_dev()).to(dtype).contiguous()


def _as_f32(t: torch.Tensor) -> np.ndarray:
    return t.float().cpu().numpy()


def _check_against_scores(scores, ids, S_ref, k, tol, id_base=0):
    scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
    B, N = S_ref.shape
    for b in range(B):
        kk = min(k, N)
        O.check_topk(scores[b], ids[b], S_ref[b], k, rel_tol=tol, id_base=id_base)
        if kk < k:
            assert np.all(ids[b, kk:] == -1) and np.all(np.isneginf(scores[b, kk:]))
        # bit-equal returned scores: exact oracle ties must come lowest id first; otherwise (two different fp32
        # keys that collide only after the final query-norm scaling) the order must agree with the oracle's
        s = scores[b, :kk]
        same = np.nonzero(s[1:] == s[:-1])[0]
        for j in same:
            a, c = ids[b, j] - id_base, ids[b, j + 1] - id_base
            assert (a < c) if S_ref[b, a] == S_ref[b, c] else (S_ref[b, a] > S_ref[b, c] - 1e-6)


def _index(X, metric="cos", **kw):
    from a_modular_rag_framework_b200 import DenseIndex
    return DenseIndex(X, metric=metric, **kw)


@pytest.mark.parametrize("name,dtype,paths", [
    ("topk_bf16_n600_d64.npz", torch.bfloat16, ("scan", "gemm")),
    ("topk_f16_n300_d384.npz", torch.float16, ("scan", "gemm")),
    ("topk_f32_n400_d48.npz", torch.float32, ("scan",)),
])
def test_golden_fixtures(golden_dir, name, dtype, paths):
    z = np.load(os.path.join(golden_dir, name))
    X, Q, S = _to_dev(z["X"], dtype), _to_dev(z["Q"], dtype), z["S"]
    assert np.array_equal(_as_f32(X), z["X"]) and np.array_equal(_as_f32(Q), z["Q"])  # fixtures are exactly representable
    ix = _index(X)
    for key in z.files:
        if not key.startswith("ids_k"):
            continue
        k = int(key[5:])
        for path in paths:
            scores, ids = ix.search(Q, k, path=path)
            _check_against_scores(scores, ids, S, k, TOL[dtype])
            # the reference's exact-tie groups (duplicated rows 3, 7, N/2, N-1) must appear lowest id first,
            # and the all-zero query must return ids 0..k-1 with score 0.0
            ids_np = ids.cpu().numpy()
            kk = min(k, X.shape[0])
            assert np.array_equal(ids_np[-1, :kk], np.arange(kk)), (path, k)
            assert np.all(scores.cpu().numpy()[-1, :kk] == 0.0)
            ref_ids = z[key]
            b = Q.shape[0] - 2  # query == row 3: its duplicates lead the list in id order
            assert np.array_equal(ids_np[b, :min(4, kk)], ref_ids[b, :min(4, kk)])


@pytest.mark.parametrize("N,D,B,k,path,dtype,metric", [
    (20000, 128, 37, 100, "gemm", torch.bfloat16, "cos"),
    (20000, 128, 3, 100, "scan", torch.bfloat16, "cos"),
    (5000, 72, 129, 10, "gemm", torch.bfloat16, "cos"),      # D % 64 != 0, two query tiles
    (70000, 64, 300, 10, "gemm", torch.float16, "ip"),
    (9000, 768, 16, 100, "gemm", torch.bfloat16, "cos"),
    (9000, 768, 16, 100, "auto", torch.bfloat16, "cos"),
    (9000, 768, 1, 100, "auto", torch.bfloat16, "cos"),
    (3001, 96, 7, 1000, "gemm", torch.bfloat16, "cos"),
    (3001, 96, 2, 1000, "scan", torch.float16, "cos"),
    (4000, 100, 5, 50, "scan", torch.float32, "cos"),
    (4000, 100, 9, 50, "scan", torch.float32, "ip"),
    (50, 64, 6, 100, "gemm", torch.bfloat16, "cos"),         # N < k
    (50, 64, 2, 100, "scan", torch.bfloat16, "cos"),
    (1, 8, 1, 1, "scan", torch.bfloat16, "cos"),
    (257, 8, 5, 3, "gemm", torch.bfloat16, "cos"),
    (30000, 64, 5, 2048, "gemm", torch.bfloat16, "cos"),     # B200R_MAX_K
    (30000, 64, 2, 2048, "scan", torch.bfloat16, "cos"),
    (40000, 128, 700, 64, "gemm", torch.float16, "cos"),     # CTA pairs, partial last query tile
    (1000, 64, 130, 5, "gemm", torch.bfloat16, "ip"),        # pair whose second CTA has 2 valid queries
    (6000, 3072, 9, 20, "gemm", torch.bfloat16, "cos"),      # text-embedding-3-large width (reference settings.yaml:8,19)
    (6000, 3072, 300, 20, "gemm", torch.bfloat16, "cos"),
    (3000, 3072, 2, 20, "scan", torch.float32, "cos"),
])
def test_random_vs_oracle(N, D, B, k, path, dtype, metric):
    rng = np.random.default_rng(N * 31 + D + B)
    X = _to_dev(rng.standard_normal((N, D)), dtype)
    Q = _to_dev(rng.standard_normal((B, D)), dtype)
    S = O.score_matrix(_as_f32(Q), _as_f32(X), metric)
    ix = _index(X, metric=metric, id_base=1000)
    scores, ids = ix.search(Q, k, path=path)
    _check_against_scores(scores, ids, S, k, TOL[dtype], id_base=1000)


@pytest.mark.parametrize("path,B", [("gemm", 40), ("scan", 4)])
def test_exact_integer_data_bit_exact_ids(path, B):
    """Small-integer vectors make every product and sum exact in fp32, so ids (ties included) must
    equal the oracle's stable-sort order exactly."""
    rng = np.random.default_rng(7)
    N, D, k = 6000, 64, 100
    X = rng.integers(-2, 3, size=(N, D)).astype(np.float32)
    Qn = rng.integers(-2, 3, size=(B, D)).astype(np.float32)
    Xd, Qd = _to_dev(X, torch.bfloat16), _to_dev(Qn, torch.bfloat16)
    ref_s, ref_i = O.topk(Qn, X, k, metric="ip")
    ix = _index(Xd, metric="ip")
    scores, ids = ix.search(Qd, k, path=path)
    assert np.array_equal(ids.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy().astype(np.float64), ref_s)


def test_constant_vectors_all_tie():
    """The Ollama mock provider returns constant vectors (reference ollama_provider.py:26-28): every
    cosine ties, so the result is ids 0..k-1."""
    X = torch.full((3000, 64), 0.1, device=_dev(), dtype=torch.bfloat16)
    Q = torch.full((5, 64), 0.1, device=_dev(), dtype=torch.bfloat16)
    ix = _index(X)
    for path in ("gemm", "scan"):
        scores, ids = ix.search(Q, 20, path=path)
        assert np.array_equal(ids.cpu().numpy(), np.tile(np.arange(20), (5, 1))), path
        assert np.allclose(scores.cpu().numpy(), 1.0, atol=1e-3)


def test_int8_corpus_scan():
    rng = np.random.default_rng(11)
    N, D, B, k = 5000, 128, 3, 20
    Xf = rng.standard_normal((N, D)).astype(np.float32)
    q8, scale = O.quantize_int8_rows(Xf)
    Qn = rng.standard_normal((B, D)).astype(np.float32)
    X = torch.from_numpy(q8).to(_dev())
    Q = torch.from_numpy(Qn).to(_dev())
    deq = q8.astype(np.float32) * scale[:, None]
    for metric in ("cos", "ip"):
        ix = _index(X, metric=metric, row_scale=torch.from_numpy(scale).to(_dev()))
        scores, ids = ix.search(Q, k)
        _check_against_scores(scores, ids, O.score_matrix(Qn, deq, metric), k, 1e-5)


@pytest.mark.parametrize("N,D,B,k,dtype", [
    (150000, 768, 1, 100, torch.float32),      # many tiles per CTA: the stage ring wraps several times
    (150000, 768, 16, 100, torch.float32),     # two passes of 8 queries
    (40001, 256, 7, 33, torch.float32),        # ragged last tile, 7 queries in an 8-wide pass
    (17, 64, 3, 5, torch.float32),             # fewer rows than one tile
    (1, 16, 1, 1, torch.float32),
    (300000, 768, 1, 100, torch.int8),
    (300000, 768, 4, 100, torch.int8),
    (65537, 384, 7, 1000, torch.int8),         # 4 + 3 queries, k = 1000
    (63, 32, 2, 10, torch.int8),
])
def test_bulk_fed_scan_shapes(N, D, B, k, dtype):
    """fp32 / int8 corpora take the bulk-copy-fed scan (scan_bulk.cu): tile / stage / pass boundaries."""
    rng = np.random.default_rng(N + 7 * D + B)
    Qn = rng.standard_normal((B, D)).astype(np.float32)
    if dtype == torch.int8:
        q8, scale = O.quantize_int8_rows(rng.standard_normal((N, D)).astype(np.float32))
        X = torch.from_numpy(q8).to(_dev())
        ref = q8.astype(np.float32) * scale[:, None]
        ix = _index(X, metric="cos", row_scale=torch.from_numpy(scale).to(_dev()))
    else:
        ref = rng.standard_normal((N, D)).astype(np.float32)
        ix = _index(torch.from_numpy(ref).to(_dev()), metric="cos")
    scores, ids = ix.search(torch.from_numpy(Qn).to(_dev()), k, path="scan")
    _check_against_scores(scores, ids, O.score_matrix(Qn, ref, "cos"), k, 1e-5)


@pytest.mark.parametrize("N,D,B,k,metric", [
    (300000, 768, 1, 100, "cos"),
    (300000, 768, 16, 100, "cos"),          # a full 16-query pass, several phases
    (65537, 384, 7, 1000, "ip"),            # ragged last tile, k = 1000
    (4000, 144, 5, 20, "cos"),              # D = 128 + 16: the second K slice is mostly out of bounds (zero-filled)
    (63, 32, 2, 10, "cos"),                 # fewer rows than one tile, D < one K slice
    (1, 16, 1, 1, "ip"),
    (20000, 256, 20, 50, "cos"),            # one group of 32 queries (96 TMEM columns), partly empty
    (9000, 128, 49, 10, "ip"),              # one group of 64 queries (192 columns, two accumulators)
    (70000, 768, 64, 100, "cos"),           # a full 64-query group at the benchmark width
    (5000, 1024, 40, 10, "cos"),            # 64 resident queries would not fit next to the ring: two groups of 32
    (3000, 64, 300, 5, "cos"),              # more than one 256-query slice
    (120000, 768, 256, 50, "cos"),          # four 64-query groups ganged in one launch (37 CTAs each), several phases
    (50000, 256, 130, 20, "ip"),            # three ganged groups, the last one mostly empty
    (40000, 1024, 200, 10, "cos"),          # seven ganged groups of 32 (21 CTAs each, one SM idle)
])
def test_int8_tensor_path(N, D, B, k, metric):
    """int8 corpora run on tcgen05 kind::i8 (rows_i8.cu), 16 / 32 / 64 queries per corpus pass: query limbs, exact int32 sums."""
    rng = np.random.default_rng(3 * N + D + B)
    Qn = rng.standard_normal((B, D)).astype(np.float32)
    q8, scale = O.quantize_int8_rows(rng.standard_normal((N, D)).astype(np.float32))
    ref = q8.astype(np.float32) * scale[:, None]
    ix = _index(torch.from_numpy(q8).to(_dev()), metric=metric, row_scale=torch.from_numpy(scale).to(_dev()), id_base=7)
    scores, ids = ix.search(torch.from_numpy(Qn).to(_dev()), k)
    _check_against_scores(scores, ids, O.score_matrix(Qn, ref, metric), k, 1e-5, id_base=7)
    s2, i2 = ix.search(torch.from_numpy(Qn).to(_dev()), k, path="gemm")
    assert torch.equal(i2, ids) and torch.equal(s2, scores)


def test_int8_tensor_scores_bit_equal_the_limb_restatement():
    """Inner-product scores of the int8 tensor path must equal, bit for bit, the CPU restatement of its arithmetic
    (oracle.int8_limb_scores: exact integer limb sums, fp32 recombination, query scale, row scale)."""
    rng = np.random.default_rng(123)
    N, D, B, k = 6000, 256, 3, 30
    q8, scale = O.quantize_int8_rows(rng.standard_normal((N, D)).astype(np.float32))
    Qn = rng.standard_normal((B, D)).astype(np.float32)
    ix = _index(torch.from_numpy(q8).to(_dev()), metric="ip", row_scale=torch.from_numpy(scale).to(_dev()))
    scores, ids = ix.search(torch.from_numpy(Qn).to(_dev()), k)
    scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
    for b in range(B):
        want = O.int8_limb_scores(Qn[b], q8, row_mult=scale)
        assert np.array_equal(scores[b], want[ids[b]])
        order = np.lexsort((np.arange(N), -want.astype(np.float64)))[:k]     # and the ranking of those exact scores
        assert np.array_equal(ids[b], order)


def test_int8_integer_data_bit_exact_ids():
    """Small-integer queries and rows: limb quantisation, int32 sums and the fp32 recombination are all exact, so ids
    must match the oracle exactly, ties (lower id first) included."""
    rng = np.random.default_rng(99)
    N, D, B, k = 30000, 64, 9, 40
    X8 = rng.integers(-3, 4, size=(N, D)).astype(np.int8)
    Qn = rng.integers(-2, 3, size=(B, D)).astype(np.float32)
    ix = _index(torch.from_numpy(X8).to(_dev()), metric="ip")
    scores, ids = ix.search(torch.from_numpy(Qn).to(_dev()), k)
    S = Qn.astype(np.float64) @ X8.astype(np.float64).T
    want_s, want_i = O.topk_from_scores(S, k)
    assert np.array_equal(ids.cpu().numpy(), want_i)
    assert np.array_equal(scores.cpu().numpy().astype(np.float64), want_s)


def test_multi_chunk_large():
    """Several units per CTA and per query (chunked rows, published thresholds)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    N, D, B, k = 300000, 64, 200, 100
    X = torch.randn((N, D), generator=g, device=_dev()).to(torch.bfloat16)
    Q = torch.randn((B, D), generator=g, device=_dev()).to(torch.bfloat16)
    ix = _index(X)
    scores, ids = ix.search(Q, k, path="gemm")
    sub = [0, 1, 99, 127, 128, 199]
    S = O.score_matrix(_as_f32(Q[sub]), _as_f32(X))
    _check_against_scores(scores[sub], ids[sub], S, k, 1e-3)
    s2, i2 = ix.search(Q[:3].contiguous(), k, path="scan")
    _check_against_scores(s2, i2, O.score_matrix(_as_f32(Q[:3]), _as_f32(X)), k, 1e-3)


def test_search_host_matches_device():
    g = torch.Generator(device="cuda").manual_seed(9)
    X = torch.randn((10000, 64), generator=g, device=_dev()).to(torch.bfloat16)
    Q = torch.randn((33, 64), generator=g, device=_dev()).to(torch.bfloat16)
    ix = _index(X)
    s_dev, i_dev = ix.search(Q, 10)
    s_h, i_h = ix.search_host(Q.cpu().pin_memory(), 10)
    assert torch.equal(i_dev.cpu(), i_h) and torch.equal(s_dev.cpu(), s_h)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_search_host_f32_queries_are_rounded_on_the_device(dtype):
    """b200r_search_host_f32 / b200r_convert_f32: fp32 provider embeddings are rounded to the index's query dtype by the
    library (RNE, bit-identical to torch's conversion, halfway cases and subnormals included), so the result equals the
    search with host-converted queries."""
    g = torch.Generator(device="cuda").manual_seed(10)
    X = torch.randn((20000, 96), generator=g, device=_dev()).to(dtype)
    Q32 = torch.randn((37, 96), generator=g, device=_dev())
    Q32[0, :8] = torch.tensor([1.00390625, 1.01171875, -1.00390625, 3.0e-39, 65504.0, 1e-8, -0.0, 0.333251953125], device=_dev())
    ix = _index(X)
    if dtype != torch.float32:
        assert torch.equal(ix.to_query_dtype(Q32.contiguous()).view(torch.int16), Q32.to(dtype).view(torch.int16))
    s_ref, i_ref = ix.search(Q32.to(dtype), 10)
    s_h, i_h = ix.search_host(Q32.cpu().pin_memory(), 10)
    assert torch.equal(i_ref.cpu(), i_h) and torch.equal(s_ref.cpu(), s_h)


def test_merge_topk_vs_oracle():
    from a_modular_rag_framework_b200 import merge_topk
    rng = np.random.default_rng(2)
    G, B, kin, k = 8, 9, 100, 100
    sc = rng.standard_normal((G, B, kin)).astype(np.float32)
    sc[:, :, 5] = 0.25  # cross-shard exact ties
    ids = np.stack([np.stack([rng.permutation(1000)[:kin] + g * 1000 for _ in range(B)]) for g in range(G)]).astype(np.int64)
    ids[3, :, 90:] = -1  # padding
    ms, mi = merge_topk(torch.from_numpy(sc).to(_dev()), torch.from_numpy(ids).to(_dev()), k)
    rs, ri = O.merge_topk(sc.transpose(1, 0, 2), ids.transpose(1, 0, 2), k)
    assert np.array_equal(mi.cpu().numpy(), ri)
    assert np.array_equal(ms.cpu().numpy(), rs)


def test_bad_arguments_raise():
    from a_modular_rag_framework_b200 import B200RError, DenseIndex
    X = torch.zeros((16, 8), device=_dev(), dtype=torch.bfloat16)
    ix = DenseIndex(X)
    with pytest.raises(B200RError):
        ix.search(torch.zeros((1, 8), device=_dev(), dtype=torch.bfloat16), 5000)
    with pytest.raises(ValueError):
        ix.search(torch.zeros((1, 8), device=_dev(), dtype=torch.float32), 5)
    with pytest.raises(ValueError):
        DenseIndex(torch.zeros((16, 8), dtype=torch.bfloat16))
    with pytest.raises(B200RError):
        DenseIndex(torch.zeros((16, 4), device=_dev(), dtype=torch.bfloat16))  # rows not 16-byte aligned


def test_min_score_floor_and_semantic_sim_edges():
    """b200r_search_min_score against the oracle, and the corpus-scale semantic_sim edge list against the
    reference's rule (edge_builder.py:146-169: every pair a<b with sim >= threshold, zero-norm -> 0.0)."""
    from a_modular_rag_framework_b200.index import semantic_sim_edges
    rng = np.random.default_rng(17)
    N, D, tau = 6000, 32, 0.55
    Xn = rng.standard_normal((N, D)).astype(np.float32)
    Xn[100:140] = Xn[100] + 0.3 * rng.standard_normal((40, D)).astype(np.float32)   # a cluster of near-duplicates
    Xn[7] = 0.0
    Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(Xn))
    X = torch.from_numpy(Xn).to(_dev()).to(torch.bfloat16)
    ix = _index(X)
    S = O.score_matrix(Xn, Xn)
    for path, B in (("gemm", 300), ("scan", 2)):
        s, i = ix.search(X[:B].contiguous(), 50, path=path, min_score=tau)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        for b in range(B):
            ref = np.nonzero(S[b] >= tau)[0]
            got = i[b][i[b] >= 0]
            sure = set(np.nonzero(S[b] >= tau + 1e-3)[0].tolist())
            maybe = set(np.nonzero(S[b] >= tau - 1e-3)[0].tolist())
            if len(ref) <= 50:
                assert sure <= set(got.tolist()) <= maybe, (path, b)
            assert np.all(s[b][: len(got)] >= tau - 1e-3) and np.all(np.isneginf(s[b][len(got):]))
    # every pair a < b with cos >= tau, no per-row cap (the cluster rows have ~40 partners each)
    got = set()
    for a0, offs, dst, sim in semantic_sim_edges(ix, tau, batch=1024):
        offs, dst, sim = offs.cpu().numpy(), dst.cpu().numpy(), sim.cpu().numpy()
        for j in range(len(offs) - 1):
            seg = dst[offs[j]:offs[j + 1]]
            assert np.all(np.diff(seg) > 0) and (len(seg) == 0 or seg[0] > a0 + j)          # ascending, strictly above the row itself
            assert np.allclose(sim[offs[j]:offs[j + 1]], S[a0 + j, seg], atol=1e-3)
            got.update((a0 + j, int(b)) for b in seg)
    iu = np.triu_indices(N, 1)
    sure = set(zip(*[a[S[iu] >= tau + 1e-3].tolist() for a in iu]))
    maybe = set(zip(*[a[S[iu] >= tau - 1e-3].tolist() for a in iu]))
    assert sure <= got <= maybe
    assert all(a < b for a, b in got) and 7 not in {a for a, _ in got} | {b for _, b in got}


@pytest.mark.parametrize("N,D,B,dtype", [(3000, 64, 3000, torch.bfloat16), (70000, 128, 5, torch.float16), (999, 40, 130, torch.bfloat16)])
def test_range_search_returns_every_hit(N, D, B, dtype):
    """b200r_range_count / b200r_range_fill: exact-size output, every row at or above the threshold, ascending ids, no cap --
    including the reference's own degenerate case (edge_builder.py:47-48: a 1-D fake embedding makes EVERY pair qualify)."""
    from a_modular_rag_framework_b200.index import range_search
    rng = np.random.default_rng(N + B)
    if N == 3000:
        Xn = np.ones((N, D), dtype=np.float32) * rng.uniform(0.5, 2.0, size=(N, 1)).astype(np.float32)   # all cosines == 1
        Xn[17] = 0.0                                                                                  # zero norm -> 0.0, never an edge
    else:
        Xn = rng.standard_normal((N, D)).astype(np.float32)
        Xn[: N // 3] = Xn[0] + 0.5 * rng.standard_normal((N // 3, D)).astype(np.float32)               # a big near-duplicate cluster
    X = _to_dev(Xn, dtype)
    Xf = _as_f32(X)
    ix = _index(X, id_base=5)
    Q = X[:B].contiguous()
    S = O.score_matrix(Xf[:B], Xf)
    tau = 0.8
    offs, ids, sc = range_search(ix, Q, tau)
    offs, ids, sc = offs.cpu().numpy(), ids.cpu().numpy() - 5, sc.cpu().numpy()
    assert offs[0] == 0 and offs[-1] == len(ids) == len(sc)
    total_sure = 0
    for b in range(B):
        seg = ids[offs[b]:offs[b + 1]]
        assert np.all(np.diff(seg) > 0)
        sure = np.nonzero(S[b] >= tau + 1e-3)[0]
        maybe = S[b] >= tau - 1e-3
        assert np.all(maybe[seg]) and np.isin(sure, seg).all()
        assert np.allclose(sc[offs[b]:offs[b + 1]], S[b, seg], atol=1e-3)
        total_sure += len(sure)
    if N == 3000:
        assert len(ids) == (N - 1) * (N - 1)                      # every non-zero row pairs with every non-zero row (itself included)
        own = torch.arange(B, device=_dev(), dtype=torch.int64) + 5
        offs2, ids2, _ = range_search(ix, Q, tau, id_gt=own)
        assert int(offs2[-1]) == (N - 1) * (N - 2) // 2            # each unordered pair once, no self pairs
        assert bool((ids2 > own.repeat_interleave(offs2[1:] - offs2[:-1])).all())
    # nothing qualifies
    o3, i3, s3 = range_search(ix, Q[:2].contiguous(), 1.5)
    assert o3.tolist() == [0, 0, 0] and i3.numel() == 0


def _device_unit_rows(n, d, seed, dtype=torch.bfloat16, block=1 << 19):
    X = torch.empty((n, d), dtype=dtype, device=_dev())
    for r0 in range(0, n, block):
        g = torch.Generator(device="cuda").manual_seed(seed + r0)
        t = torch.randn((min(block, n - r0), d), generator=g, device=_dev(), dtype=torch.float32)
        t /= t.norm(dim=1, keepdim=True)
        X[r0:r0 + t.shape[0]] = t.to(dtype)
    return X


@pytest.mark.parametrize("N,D,B,k,nq", [
    (1_000_000, 768, 1024, 100, 24),     # BASELINE.json configs[1] at full size
    (2_100_000, 768, 640, 100, 16),      # 5 phases, 37 chunks, CTA pairs with a half-empty last pair (the 10M x 4096 schedule's shape)
    (1_250_000, 768, 4096, 100, 16),     # one rank's shard of configs[2] at N=8, full batch
])
def test_benchmark_scale_configs_vs_oracle(N, D, B, k, nq):
    """The configurations the throughput numbers are quoted on, checked like bench.py's recall leg: fp64 oracle top-k of
    `nq` queries over ALL rows (streamed back in blocks), tie-aware rule, 1e-3 relative scores; plus exact ties planted
    across distant tiles / phases (duplicates of row 17), which must come out adjacent, lowest id first."""
    X = _device_unit_rows(N, D, seed=N + B)
    dup = [17, 1000, N // 3, N - 5]
    for r in dup[1:]:
        X[r] = X[17]
    g = torch.Generator(device="cuda").manual_seed(99)
    Q = torch.randn((B, D), generator=g, device=_dev(), dtype=torch.float32)
    Q /= Q.norm(dim=1, keepdim=True)
    Q = Q.to(torch.bfloat16)
    Q[5] = X[17]
    ix = _index(X)
    scores, ids = ix.search(Q, k)
    sel = np.unique(np.concatenate([[5, 0, B - 1, 127, 128, 255, 256], np.linspace(0, B - 1, nq).astype(np.int64)]))
    Qs = Q[torch.from_numpy(sel).to(_dev())].float().cpu().numpy()

    def blocks(rows=1 << 18):
        for r0 in range(0, N, rows):
            yield r0, O.bf16_bits_to_f32(X[r0:r0 + rows].view(torch.int16).cpu().numpy().view(np.uint16))
    ref_s, ref_i = O.topk_blocks_exact(Qs, blocks(), k, extra=64)
    gs, gi = scores.cpu().numpy(), ids.cpu().numpy()
    for j, b in enumerate(sel):
        r = O.check_topk_against_list(gs[b], gi[b], ref_s[j], ref_i[j], k, 1e-3)
        assert r["ok"] and r["recall"] == 1.0, (int(b), r)
    assert gi[5, :4].tolist() == sorted(dup)
    assert np.all(gs[5, :4] == gs[5, 0])
    # whole-batch invariants at full size: ids unique per query, scores sorted
    assert bool((scores[:, :-1] >= scores[:, 1:]).all())
    srt = ids.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())


def test_search_captured_in_a_cuda_graph_replays():
    """A whole search step (init, tile phases, selects) captured once and replayed with new queries in the static buffer."""
    from a_modular_rag_framework_b200.index import CapturedSearch
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn((200000, 64), generator=g, device=_dev()).to(torch.bfloat16)
    ix = _index(X)
    for B in (1, 16, 300):
        Qs = [torch.randn((B, 64), generator=g, device=_dev()).to(torch.bfloat16) for _ in range(3)]
        cap = CapturedSearch(ix, Qs[0], 50)
        for q in Qs:
            s, i = cap(q)
            rs, ri = ix.search(q, 50)
            assert torch.equal(i, ri) and torch.equal(s, rs)


@pytest.mark.parametrize("N,D,B,k,metric", [
    (150000, 768, 1, 100, "cos"),
    (150000, 768, 16, 100, "cos"),          # a full 16-query pass: the corpus is streamed once
    (300000, 384, 9, 100, "ip"),            # several phases
    (40001, 256, 7, 33, "cos"),             # ragged last tile
    (5000, 100, 5, 50, "cos"),              # D % 32 != 0: the last K slice is partly out of bounds (zero-filled)
    (17, 64, 3, 5, "cos"),                  # fewer rows than one tile
    (1, 16, 1, 1, "ip"),
    (60000, 768, 40, 10, "cos"),            # three groups of 16 queries ganged in one launch (49 CTAs each)
    (30000, 1024, 4, 20, "cos"),
    (100000, 768, 256, 20, "cos"),          # a full 256-query slice: 16 ganged groups, 9 CTAs each
    (20000, 128, 300, 10, "ip"),            # two slices, the second one with 3 groups
])
def test_fp32_tensor_path(N, D, B, k, metric):
    """fp32 corpora on tcgen05 kind::tf32 with both operands split into two exact tf32 pieces (rows_f32.cu): the fp32 bar,
    1e-5 relative against the fp64 oracle, on data whose components span many binades (so the low mantissa bits matter)."""
    rng = np.random.default_rng(5 * N + D + B)
    Xn = (rng.standard_normal((N, D)) * np.exp2(rng.integers(-6, 7, size=(N, D)))).astype(np.float32)
    Qn = (rng.standard_normal((B, D)) * np.exp2(rng.integers(-6, 7, size=(B, D)))).astype(np.float32)
    if N > 100:
        Xn[N // 2] = Xn[3]                  # exact duplicate: ties -> lower id first
        Qn[0] = Xn[3]
    X, Q = torch.from_numpy(Xn).to(_dev()), torch.from_numpy(Qn).to(_dev())
    ix = _index(X, metric=metric, id_base=11)
    S = O.score_matrix(Qn, Xn, metric)
    scores, ids = ix.search(Q, k, path="gemm")
    _check_against_scores(scores, ids, S, k, 1e-5, id_base=11)
    s_auto, i_auto = ix.search(Q, k)                       # auto == the tensor path for fp32 corpora it fits
    assert torch.equal(i_auto, ids) and torch.equal(s_auto, scores)
    if N > 100:
        assert ids[0, :2].tolist() == [3 + 11, N // 2 + 11]
    # the CUDA-core scan agrees to the same bar
    s2, i2 = ix.search(Q, k, path="scan")
    _check_against_scores(s2, i2, S, k, 1e-5, id_base=11)


def test_fp32_split_is_tight_on_nearly_parallel_vectors():
    """Cosines packed within 1e-6 of each other near 1.0: a plain tf32 product (2^-11 per operand) would scramble them; the
    split-precision path must rank them like the fp64 oracle wherever they differ by more than the fp32 bar."""
    rng = np.random.default_rng(1)
    N, D, k = 20000, 768, 50
    base = rng.standard_normal(D).astype(np.float32)
    Xn = base[None, :] * (1.0 + 1e-3 * rng.standard_normal((N, D)).astype(np.float32))
    Qn = base[None, :].copy()
    ix = _index(torch.from_numpy(Xn).to(_dev()), metric="cos")
    scores, ids = ix.search(torch.from_numpy(Qn).to(_dev()), k, path="gemm")
    S = O.score_matrix(Qn, Xn)
    got = scores.cpu().numpy()[0].astype(np.float64)
    # far inside 1e-5: the split loses < 2^-22 per operand and the accumulation runs in 8 short chains (a single chain of
    # truncating tensor-core adds measured 1.5e-5 here)
    assert np.max(np.abs(got - S[0, ids.cpu().numpy()[0]])) < 4e-6
    _check_against_scores(scores, ids, S, k, 1e-5)


def test_mmr_diversify_golden_and_random(golden_dir):
    """b200r_mmr_diversify against picks recorded from the reference (mmr_cases.json: corpus rows = the item vectors) and
    against the oracle restatement on random clustered candidates for a whole batch."""
    import json
    from a_modular_rag_framework_b200.index import mmr_diversify
    with open(os.path.join(golden_dir, "mmr_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        X = O.bf16_bits_to_f32(np.array(c["X_bf16"], dtype=np.uint16))
        n = len(c["ids"])
        ix = _index(_to_dev(X, torch.bfloat16), metric="cos", id_base=100)
        ids = torch.tensor([[(100 + j) if j != c["none"] else 100 + n + 5 for j in range(n)]], device=_dev())   # the None item: an id with no row
        sc = torch.tensor([c["scores"]], device=_dev(), dtype=torch.float32)
        oi, osc, pos = mmr_diversify(ix, ids, sc, c["top_k"], c["lambda"])
        want = [c["ids"].index(p) for p in c["picked"]]
        got = pos[0].cpu().tolist()
        assert got[:len(want)] == want and all(g == -1 for g in got[len(want):]), (c["top_k"], c["lambda"])
        assert oi[0, :len(want)].cpu().tolist() == [int(ids[0, w]) for w in want]
    # a batch of random clustered lists from a larger corpus, with padding
    rng = np.random.default_rng(9)
    N, D, B, P, top_k = 5000, 64, 7, 60, 12
    centers = rng.standard_normal((20, D)).astype(np.float32)
    Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(centers[rng.integers(0, 20, N)] + 0.2 * rng.standard_normal((N, D)).astype(np.float32)))
    ix = _index(_to_dev(Xn, torch.bfloat16))
    ids = np.stack([rng.permutation(N)[:P] for _ in range(B)]).astype(np.int64)
    ids[2, 40:] = -1
    sc = rng.uniform(0, 1, size=(B, P)).astype(np.float32)
    oi, osc, pos = mmr_diversify(ix, torch.from_numpy(ids).to(_dev()), torch.from_numpy(sc).to(_dev()), top_k, 0.6)
    for b in range(B):
        items = [(int(i), float(s), Xn[i].astype(float).tolist()) for i, s in zip(ids[b], sc[b]) if i >= 0]
        want = [p[0] for p in O.mmr_diversify(items, top_k=top_k, lambda_weight=0.6)]
        got = oi[b].cpu().tolist()
        if got[:len(want)] != want:          # fp32 vs fp64 cosines: only picks whose MMR values tie within 1e-5 may swap
            first = next(j for j in range(len(want)) if got[j] != want[j])
            assert False, (b, first, got, want)
