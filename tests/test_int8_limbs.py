"""CPU checks of the int8 paths' query arithmetic (oracle.query_limbs_int8 restates csrc/rows_i8.cu / scan_bulk.cu):
the limb split is exact, the fixed-point image of a query is within 2^-22 of its largest component, and the scores it
produces agree with the fp64 metric of retrieval_backend.py:192-197 far inside the 1e-5 bar the GPU tests use."""
import numpy as np
import pytest

from oracle import retrieval_oracle as O


@pytest.mark.parametrize("D,seed", [(16, 0), (384, 1), (768, 2), (3072, 3)])
def test_limb_split_is_exact_and_tight(D, seed):
    rng = np.random.default_rng(seed)
    q = (rng.standard_normal(D) * rng.choice([1e-6, 1.0, 300.0])).astype(np.float32)
    limbs, scale = O.query_limbs_int8(q)
    qi = limbs[2].astype(np.int64) * 65536 + limbs[1].astype(np.int64) * 256 + limbs[0].astype(np.int64)
    assert np.abs(qi).max() <= 2 ** 22
    assert np.abs(limbs[2].astype(np.int64)).max() <= 64
    qmax = float(np.abs(q).max())
    assert np.all(np.abs(q.astype(np.float64) - scale * qi) <= 2.0 ** -22 * qmax)


def test_zero_and_integer_queries():
    limbs, scale = O.query_limbs_int8(np.zeros(32, dtype=np.float32))
    assert scale == 0.0 and not limbs.any()
    q = np.array([2, -1, 0, 1, -2, 2, 0, 1], dtype=np.float32)
    limbs, scale = O.query_limbs_int8(q)
    qi = limbs[2].astype(np.int64) * 65536 + limbs[1].astype(np.int64) * 256 + limbs[0].astype(np.int64)
    assert np.array_equal(qi * 2 ** -21, q.astype(np.float64))       # exact: 2 * 2^-22 per unit
    X8 = np.arange(-8, 8, dtype=np.int8).reshape(2, 8)
    assert np.array_equal(O.int8_limb_scores(q, X8).astype(np.float64), X8.astype(np.float64) @ q.astype(np.float64))


@pytest.mark.parametrize("N,D", [(2000, 128), (500, 768)])
def test_scores_agree_with_the_fp64_metric(N, D):
    rng = np.random.default_rng(N + D)
    q8, scale = O.quantize_int8_rows(rng.standard_normal((N, D)).astype(np.float32))
    deq = q8.astype(np.float32) * scale[:, None]
    for b in range(4):
        q = rng.standard_normal(D).astype(np.float32)
        ref = O.score_matrix(q[None, :], deq, "ip")[0]
        got = O.int8_limb_scores(q, q8, row_mult=scale).astype(np.float64)
        assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))
