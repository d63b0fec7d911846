"""BASELINE.json configs[4] on N GPUs (torchrun, one rank per GPU): graph-enhanced 2-hop over a ROW-SHARDED corpus.
top-20 seed passages (sharded dense search) -> CSR expansion over a synthetic 5M-node / 50M-directed-edge graph
(replicated) -> owner-computes re-score -> min/max all-reduce -> fusion + per-rank top-10 -> all-gather + merge.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_graph_sharded.py [--batch 4096] [--steps 20]

Rank 0 prints one JSON line: whole-job queries/s (max over ranks of the device time), per-rank gather-kernel time."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=5_000_000)
    ap.add_argument("--edges", type=int, default=25_000_000, help="undirected edges (x2 directed)")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--streams", type=int, default=1, help="batches in flight (round-robin over this many CUDA streams / stage objects): "
                    "the collectives and latency-bound kernels of one batch overlap the gather of the next")
    a = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, _native
    from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, ShardedGraphStage, shard_rows

    V, D, B = a.nodes, a.dim, a.batch
    lo, hi = shard_rows(V, world, rank)
    X = torch.empty((hi - lo, D), dtype=torch.bfloat16, device=dev)
    blk = 1 << 20
    for r0 in range(lo, hi, blk):
        r1 = min(hi, r0 + blk)
        g = torch.Generator(device=dev).manual_seed(1_000_003 + r0)
        t = torch.randn((r1 - r0, D), generator=g, device=dev)
        t /= t.norm(dim=1, keepdim=True)
        X[r0 - lo:r1 - lo] = t.to(torch.bfloat16)
    # the same graph on every rank (replicated CSR)
    g = torch.Generator(device=dev).manual_seed(5)
    src = torch.randint(0, V, (a.edges,), generator=g, device=dev)
    dst = torch.randint(0, V, (a.edges,), generator=g, device=dev)
    s2, d2 = torch.cat([src, dst]), torch.cat([dst, src])
    order = torch.argsort(s2 * V + d2)
    s2, d2 = s2[order], d2[order]
    row_ptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(s2, minlength=V), 0)
    col = d2.to(torch.int32).contiguous()
    del src, dst, s2, d2, order
    graph = PassageGraph(row_ptr, col)
    local = DenseIndex(X, metric="cos", id_base=lo)
    nb = 4
    gq = torch.Generator(device=dev).manual_seed(99)
    Q = torch.randn((nb, B, D), generator=gq, device=dev)
    Q = (Q / Q.norm(dim=2, keepdim=True)).to(torch.bfloat16)          # same queries on every rank
    if world > 1:
        sidx = ShardedDenseIndex(local, exchange="nccl")
        seeds = [sidx.search(Q[i], a.seeds)[1].contiguous() for i in range(nb)]
    else:
        seeds = [local.search(Q[i], a.seeds)[1].contiguous() for i in range(nb)]
    stages = [ShardedGraphStage(local, graph, V) for _ in range(a.streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(a.streams)] if a.streams > 1 else [torch.cuda.current_stream(dev)]

    def step(i):
        j = i % a.streams
        if a.streams == 1:
            return stages[0].run(Q[i % nb], seeds[i % nb], 2, 0.4, 0.2, a.k)
        with torch.cuda.stream(streams[j]):
            return stages[j].run(Q[i % nb], seeds[i % nb], 2, 0.4, 0.2, a.k)

    def join():
        cur = torch.cuda.current_stream(dev)
        for s_ in streams:
            if s_ is not cur:
                cur.wait_stream(s_)

    for i in range(a.warmup):
        out = step(i)
    join()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.streams > 1:
        for s_ in streams:
            s_.wait_stream(torch.cuda.current_stream(dev))
    e0.record()
    for i in range(a.steps):
        out = step(i)
    join()
    e1.record()
    torch.cuda.synchronize()
    gms, gn = _native.profile_read()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / a.steps
    per_rank = {"rank": rank, "ms_per_batch": ms, "gather_kernel_ms": gms / max(gn, 1)}
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        allr = [None] * world
        dist.all_gather_object(allr, per_rank)
    else:
        allr = [per_rank]
    if rank == 0:
        ncand = out["ncand"].float().mean().item()
        print(json.dumps({"config": "cfg4_graph_sharded", "n_gpus": world, "nodes": V, "edges_directed": 2 * a.edges, "dim": D, "batch": B,
                          "seeds": a.seeds, "k": a.k, "window": 2, "batches_in_flight": a.streams, "ms_per_batch": ms, "queries_per_s": B / (ms * 1e-3),
                          "avg_candidates": ncand, "overflowed_queries": int((out["status"] != 0).sum().item()), "per_rank": allr}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
