"""World-size-2 (and 3) CPU test of the sharded search: row partition, global ids, packed all-gather,
merge order.  The local engine and the merge are oracle stand-ins here (test infrastructure); the GPU
path is covered by tests/test_gpu_sharded.py on the B200 box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval_oracle as O
from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, shard_rows


def test_shard_rows_cover_exactly():
    for N in (1, 7, 100, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(N, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(lo <= hi for lo, hi in spans)


def test_exchange_auto_picks_collectives_without_a_cuda_index():
    """exchange="auto": the library's peer-memory exchange needs a CUDA DenseIndex; any other local engine (these CPU
    stand-ins) gets the collective path, and an unknown name is rejected."""
    local = type("L", (), {"search": lambda *a, **k: None})()
    assert ShardedDenseIndex(local, merge_packed=lambda *a: None).exchange == "nccl"
    cuda_like = type("C", (), {"search_sharded_begin": lambda *a, **k: None, "device": torch.device("cuda", 0)})()
    assert ShardedDenseIndex(cuda_like).exchange == "abi"
    assert ShardedDenseIndex(cuda_like, exchange="p2p").exchange == "abi"          # round-1 name
    with pytest.raises(ValueError):
        ShardedDenseIndex(local, exchange="smoke-signals")


class _OracleLocal:
    def __init__(self, X, id_base):
        self.X, self.id_base = X, id_base

    def search(self, Q, k, path="auto", out=None):
        if self.X.shape[0] == 0:
            s = np.full((Q.shape[0], k), -np.inf, dtype=np.float32)
            i = np.full((Q.shape[0], k), -1, dtype=np.int64)
        else:
            s, i = O.topk(Q.numpy(), self.X, k, id_base=self.id_base)
            s = np.pad(s.astype(np.float32), ((0, 0), (0, k - s.shape[1])), constant_values=-np.inf)
            i = np.pad(np.where(i >= self.id_base, i, -1), ((0, 0), (0, k - i.shape[1])), constant_values=-1)
        out[0].copy_(torch.from_numpy(s))
        out[1].copy_(torch.from_numpy(i))
        return out


def _oracle_merge_packed(packed, G, B, k):
    n = B * k
    rec = packed.view(G, 12 * n)
    ids = torch.stack([rec[g, : 8 * n].view(torch.int64).view(B, k) for g in range(G)], 1).numpy()
    sc = torch.stack([rec[g, 8 * n:].view(torch.float32).view(B, k) for g in range(G)], 1).numpy()
    ms, mi = O.merge_topk(sc, ids, k)
    return torch.from_numpy(ms), torch.from_numpy(mi)


def _worker(rank, world, port, N, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        X = rng.standard_normal((N, 16)).astype(np.float32)
        X[N - 1] = X[0]          # a tie that straddles shards
        X[N // 2] = X[0]
        Q = rng.standard_normal((4, 16)).astype(np.float32)
        Q[1] = X[0]
        lo, hi = shard_rows(N, world, rank)
        sh = ShardedDenseIndex(_OracleLocal(X[lo:hi], lo), merge_packed=_oracle_merge_packed)
        s, i = sh.search(torch.from_numpy(Q), k)
        fs, fi = O.topk(Q, X, k)
        ok = np.array_equal(i.numpy()[:, : fi.shape[1]], fi) and np.allclose(s.numpy()[:, : fs.shape[1]], fs, atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,N,k", [(2, 1001, 10), (3, 50, 20), (2, 3, 4)])
def test_sharded_equals_unsharded_gloo(world, N, k):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
