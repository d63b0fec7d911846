"""Sharded search on the GPU: virtual shards on one device (packed merge kernel + id_base), and a real
2-rank NCCL run when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as O
from tests.helpers_graph import check_graph_query

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
        sh = ShardedDenseIndex(DenseIndex(X, id_base=lo), exchange="nccl")      # the collective fallback first
        assert ShardedDenseIndex(sh.local).exchange == "abi"                    # what "auto" picks for a CUDA index
        s, i = sh.search(torch.from_numpy(Qn).to(dev).to(torch.bfloat16), k)
        S = O.score_matrix(Qn, Xn)
        for b in range(B):
            O.check_topk(s[b].cpu().numpy(), i[b].cpu().numpy(), S[b], k, rel_tol=1e-3)
        # overlapped exchange (side stream, double-buffered): several batches in flight, identical results
        Qb = torch.from_numpy(Qn).to(dev).to(torch.bfloat16)
        pend = [sh.search(Qb, k, overlap=True) for _ in range(5)]
        for pr in pend:
            s2, i2 = pr.get()
            assert torch.equal(i2, i) and torch.equal(s2, s)
        # C-ABI exchange (peer stores + device-side flags, csrc/comm.cu); descriptors swapped through torch.distributed
        shp = ShardedDenseIndex(sh.local, exchange="abi")
        for _ in range(3):
            s3, i3 = shp.search(Qb, k)
            assert torch.equal(i3, i) and torch.equal(s3, s)
        pend = [shp.search(Qb, k, overlap=True) for _ in range(5)]
        for pr in pend:
            s3, i3 = pr.get()
            assert torch.equal(i3, i) and torch.equal(s3, s)
        shp.comm.check()
        # graph stage over the same row shards: CSR replicated, owner computes, min/max all-reduced
        from a_modular_rag_framework_b200 import PassageGraph
        from a_modular_rag_framework_b200.sharded import ShardedGraphStage
        row_ptr, col_idx = O.csr_from_edges(N, rng.integers(0, N, 4 * N), rng.integers(0, N, 4 * N))
        graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
        stage = ShardedGraphStage(sh.local, graph, N)
        seeds = i[:, :8].contiguous()
        out = stage.run(torch.from_numpy(Qn).to(dev).to(torch.bfloat16), seeds, 2, 0.4, 0.2, 10)
        for b in range(0, B, 3):
            check_graph_query(out, b, Qn[b], Xn, row_ptr, col_idx, seeds[b].cpu().numpy(), 2, 0.4, 0.2, 10)
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


def test_graph_stage_virtual_shards_equal_single_gpu():
    """Owner-computes graph stage: G virtual ranks on one device (pass 1 per shard, min/max combined like the
    all-reduce, pass 2 per shard, merge) must reproduce the single-launch result."""
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore
    from a_modular_rag_framework_b200.index import ExpandShardState, expand_rescore_pass
    from a_modular_rag_framework_b200.sharded import owned_range, shard_rows
    from tests.helpers_sharded import combine_shard_records
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    V, E, D, B, S, k, G = 30000, 120000, 64, 16, 20, 10, 4
    row_ptr, col_idx = O.csr_from_edges(V, rng.integers(0, V, E), rng.integers(0, V, E))
    Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((V - 500, D)).astype(np.float32)))  # last 500 nodes have no row
    Qn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((B, D)).astype(np.float32)))
    seeds = torch.from_numpy(rng.integers(0, V, size=(B, S))).to(dev)
    X = torch.from_numpy(Xn).to(dev).to(torch.bfloat16)
    Q = torch.from_numpy(Qn).to(dev).to(torch.bfloat16)
    graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
    ref = expand_rescore(DenseIndex(X), graph, Q, seeds, 2, 0.4, 0.2, k)
    N = X.shape[0]
    shards, states = [], []
    for r in range(G):
        lo, hi = shard_rows(N, G, r)
        ix = DenseIndex(X[lo:hi], id_base=lo)
        st = ExpandShardState(ix, B, k)
        expand_rescore_pass(ix, graph, Q, seeds, 2, 0.4, 0.2, k, st, 1, *owned_range(N, G, r))
        shards.append(ix); states.append(st)
    mn = torch.stack([s.minmax[:, 0] for s in states]).min(0).values
    mx = torch.stack([s.minmax[:, 1] for s in states]).max(0).values
    for r, (ix, st) in enumerate(zip(shards, states)):
        st.minmax[:, 0], st.minmax[:, 1] = mn, mx
        expand_rescore_pass(ix, graph, Q, seeds, 2, 0.4, 0.2, k, st, 2, *owned_range(N, G, r))
    # what the ranks' single all-gather of the packed records produces, viewed per field (shard stride = one record)
    gathered = torch.stack([s.record for s in states])
    out = combine_shard_records(states[0].views(gathered), states[0].out["ncand"], states[0].out["status"], k)
    assert torch.equal(out["ncand"], ref["ncand"])
    assert torch.equal(out["id"], ref["id"])
    assert torch.allclose(out["score"], ref["score"], atol=1e-6)
    assert torch.equal(out["hop"], ref["hop"])
    assert torch.allclose(out["score_dense_norm"], ref["score_dense_norm"], atol=1e-6)
    assert torch.allclose(out["cos"], ref["cos"], atol=0)
    # and against the oracle
    for b in (0, 7, 15):
        check_graph_query(out, b, Qn[b], Xn, row_ptr, col_idx, seeds[b].cpu().numpy(), 2, 0.4, 0.2, k)


def test_graph_stage_ranks_that_own_no_candidates():
    """All seeds and their neighbourhoods live in the first shard's row block: the other virtual ranks own nothing, report
    (+inf, -inf) cosine ranges and all-padding lists, and the merge must still equal the single-launch result."""
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore
    from a_modular_rag_framework_b200.index import ExpandShardState, expand_rescore_pass
    from a_modular_rag_framework_b200.sharded import owned_range, shard_rows
    from tests.helpers_sharded import combine_shard_records
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    V, D, B, S, k, G = 9000, 64, 5, 6, 8, 3
    src, dst = rng.integers(0, 2000, 8000), rng.integers(0, 2000, 8000)       # edges only among nodes < 2000 (shard 0: rows < 3000)
    row_ptr, col_idx = O.csr_from_edges(V, src, dst)
    X = torch.from_numpy(rng.standard_normal((V, D)).astype(np.float32)).to(dev).to(torch.bfloat16)
    Q = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)).to(dev).to(torch.bfloat16)
    seeds = torch.from_numpy(rng.integers(0, 2000, size=(B, S))).to(dev)
    graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
    ref = expand_rescore(DenseIndex(X), graph, Q, seeds, 2, 0.4, 0.2, k)
    shards, states = [], []
    for r in range(G):
        lo, hi = shard_rows(V, G, r)
        ix = DenseIndex(X[lo:hi], id_base=lo)
        st = ExpandShardState(ix, B, k)
        expand_rescore_pass(ix, graph, Q, seeds, 2, 0.4, 0.2, k, st, 1, *owned_range(V, G, r))
        shards.append(ix); states.append(st)
    assert torch.isinf(states[1].minmax).all() and torch.isinf(states[2].minmax).all()
    mn = torch.stack([s.minmax[:, 0] for s in states]).min(0).values
    mx = torch.stack([s.minmax[:, 1] for s in states]).max(0).values
    for r, (ix, st) in enumerate(zip(shards, states)):
        st.minmax[:, 0], st.minmax[:, 1] = mn, mx
        expand_rescore_pass(ix, graph, Q, seeds, 2, 0.4, 0.2, k, st, 2, *owned_range(V, G, r))
    assert (states[1].out["id"] == -1).all() and (states[2].out["id"] == -1).all()
    out = combine_shard_records(states[0].views(torch.stack([s.record for s in states])), states[0].out["ncand"], states[0].out["status"], k)
    for key in ("id", "hop", "ncand"):
        assert torch.equal(out[key], ref[key])
    assert torch.allclose(out["score"], ref["score"], atol=1e-6)
    assert torch.equal(out["cos"], ref["cos"])
