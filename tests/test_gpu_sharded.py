"""Sharded search on the GPU: virtual shards on one device (packed merge kernel + id_base), and a real
2-rank NCCL run when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as O

pytestmark = pytest.mark.gpu


def test_virtual_shards_equal_unsharded():
    from a_modular_rag_framework_b200 import DenseIndex
    from a_modular_rag_framework_b200.sharded import _cuda_merge_packed, shard_rows
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(3)
    N, D, B, k, G = 50001, 64, 24, 100, 4
    X = torch.randn((N, D), generator=g, device=dev).to(torch.bfloat16)
    X[40000] = X[7]                    # exact tie across shards: lower GLOBAL id must come first
    Q = torch.randn((B, D), generator=g, device=dev).to(torch.bfloat16)
    Q[0] = X[7]
    full_s, full_i = DenseIndex(X).search(Q, k)
    n = B * k
    gather = torch.empty(G * 12 * n, dtype=torch.uint8, device=dev)
    keep = []
    for r in range(G):
        lo, hi = shard_rows(N, G, r)
        ix = DenseIndex(X[lo:hi], id_base=lo)
        keep.append(ix)
        rec = gather[r * 12 * n:(r + 1) * 12 * n]
        ix.search(Q, k, out=(rec[8 * n:].view(torch.float32).view(B, k), rec[:8 * n].view(torch.int64).view(B, k)))
    ms, mi = _cuda_merge_packed(gather, G, B, k)
    assert torch.equal(mi, full_i)
    assert torch.equal(ms, full_s)
    pos7, pos40k = (mi[0] == 7).nonzero().item(), (mi[0] == 40000).nonzero().item()
    assert pos40k == pos7 + 1
    S = O.score_matrix(Q[:2].float().cpu().numpy(), X.float().cpu().numpy())
    for b in range(2):
        O.check_topk(ms[b].cpu().numpy(), mi[b].cpu().numpy(), S[b], k, rel_tol=1e-3)


def _nccl_worker(rank, world, port, ret):
    import torch.distributed as dist
    from a_modular_rag_framework_b200 import DenseIndex
    from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rng = np.random.default_rng(77)
        N, D, B, k = 30000, 64, 10, 20
        Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((N, D)).astype(np.float32)))
        Qn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((B, D)).astype(np.float32)))
        lo, hi = shard_rows(N, world, rank)
        X = torch.from_numpy(Xn[lo:hi]).to(dev).to(torch.bfloat16)
        sh = ShardedDenseIndex(DenseIndex(X, id_base=lo))
        s, i = sh.search(torch.from_numpy(Qn).to(dev).to(torch.bfloat16), k)
        S = O.score_matrix(Qn, Xn)
        for b in range(B):
            O.check_topk(s[b].cpu().numpy(), i[b].cpu().numpy(), S[b], k, rel_tol=1e-3)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_search():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)
