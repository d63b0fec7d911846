#!/usr/bin/env python3
"""bench.py -- the retrieval hot path on B200: exact top-k queries/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic queries: exact cosine top-100 of
B=4096 bf16 queries against the resident 10M x 768 bf16 passage-embedding corpus (BASELINE.json
configs[2], the configuration the metric is quoted on; it fits one GPU).  With N > 1 (launched by
torchrun, one rank per GPU) the corpus is row-sharded, every rank searches its rows, the packed local
top-k lists are all-gathered over NCCL and merged on the device: strong scaling.

After the timed regions a RECALL leg checks the results those regions produced (device-resident and e2e) against the
fp64 oracle over the whole (sharded) corpus, and the run exits non-zero if recall < 1 under the tie-aware rule; short
secondary legs ("legs": batch-1 / batch-16 latency on the same shards, the graph stage of configs[4]) put the other
north-star targets into the same line.

Printed JSON line (rank 0): value = whole-job queries/s with the queries already in HBM;
e2e = the same metric through the host-buffer C-ABI call (pinned host queries in, host results out,
copies inside the timed region); roofline = the dominant kernel (tcgen05 GEMM + fused top-k) against
the measured dense-bf16 peak; cpu_baseline = the oracle (numpy/BLAS restatement of the reference's
cosine + stable sort) timed on the host cores over a bounded sample.

--impl reference times that CPU restatement alone (the reference itself is pure Python and is not
present on the GPU box; the oracle port is its stand-in), on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exact top-k queries/sec @10Mx768 bf16"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--path", default="auto", choices=["auto", "gemm", "scan"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"], help="corpus / query element type (the contract line is bf16)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-plugin-leg", action="store_true", help="N>1: skip the leg where ONE process drives all N GPUs through the plug-in engine")
    ap.add_argument("--no-legs", action="store_true", help="skip the secondary legs (batch 1 / batch 16 latency, graph stage)")
    ap.add_argument("--graph-nodes", type=int, default=5_000_000, help="graph leg: nodes (= rows of its corpus, taken from the shards)")
    ap.add_argument("--graph-edges", type=int, default=25_000_000, help="graph leg: directed edges before symmetrisation")
    ap.add_argument("--verify-queries", type=int, default=8,
                    help="queries of the last timed batch checked against the CPU oracle after the timed regions (0 = skip)")
    ap.add_argument("--exchange", default=os.environ.get("B200R_EXCHANGE", "abi"), choices=["abi", "nccl"],
                    help="N>1: the library's own exchange over peer memory (b200r_comm_*, default) or a packed NCCL all-gather")
    return ap.parse_args()


def workload_name(a):
    return f"synthetic unit-norm {a.rows}x{a.dim} {a.dtype} corpus, batch {a.batch}, exact cosine top-{a.k}"


def shared_config(a):
    """The `config` of BOTH arms (the driver compares them): the workload only; how each arm runs it is in `detail`."""
    return {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
            "l2": "inputs (corpus >= 1.9 GB per GPU) are far larger than the 126 MB L2; no flush between steps"}


# ------------------------------------------------------------------------------------------------
# CPU leg (oracle port).  Only this function touches oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_qps(rows, dim, batch, k, target_seconds, steps=1, warmup=0, seed=1234):
    """Oracle (reference-semantics numpy restatement, float32 BLAS, all host threads) on a bounded
    sample: Bs queries x Ns rows, timed, then scaled linearly in N to the full corpus."""
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = str(cores)
    import numpy as np
    from oracle import retrieval_oracle as O
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:  # noqa: BLE001
        pass
    rng = np.random.default_rng(seed)
    Bs = min(batch, 256)
    # calibrate on a small block, then size the sample for ~target_seconds
    cal_rows = min(rows, 50_000)
    Xc = rng.standard_normal((cal_rows, dim), dtype=np.float32)
    Xc /= np.linalg.norm(Xc, axis=1, keepdims=True)
    Xc = O.bf16_bits_to_f32(O.f32_to_bf16_bits(Xc))
    Q = rng.standard_normal((Bs, dim), dtype=np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    Q = O.bf16_bits_to_f32(O.f32_to_bf16_bits(Q))
    t0 = time.perf_counter()
    O.topk(Q, Xc, min(k, cal_rows), dtype=np.float32, block_rows=1 << 16)
    t_cal = max(time.perf_counter() - t0, 1e-4)
    per_step = target_seconds / max(1, steps + warmup)
    Ns = int(min(rows, max(cal_rows, cal_rows * per_step / t_cal)))
    Ns = min(Ns, 2_000_000)
    if Ns > cal_rows:
        extra = rng.standard_normal((Ns - cal_rows, dim), dtype=np.float32)
        extra /= np.linalg.norm(extra, axis=1, keepdims=True)
        X = np.concatenate([Xc, O.bf16_bits_to_f32(O.f32_to_bf16_bits(extra))])
        del extra
    else:
        X = Xc[:Ns]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.topk(Q, X, min(k, Ns), dtype=np.float32, block_rows=1 << 16)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    qps_sample = Bs / t
    qps_full = qps_sample * (Ns / rows)
    sample = (f"{Bs} queries x {Ns} rows x {dim} (float32 BLAS + argpartition + exact merge, oracle port), "
              f"{t:.2f} s/step; queries/s scaled by {Ns}/{rows} rows to the full corpus")
    return qps_full, cores, sample, t


def cpu_python_style_qps(rows, dim, k, target_seconds=3.0, seed=77):
    """The reference's own algorithm SHAPE on one core: DenseReranker._cosine restated line by line (three generator
    passes per pair in the interpreter, retrieval_backend.py:192-197) + the stable sort/slice (:371-372), timed on a few
    thousand rows and extrapolated linearly to the full corpus.  Reported next to the (much faster) numpy/BLAS port."""
    import numpy as np
    from oracle import retrieval_oracle as O
    rng = np.random.default_rng(seed)
    q = rng.standard_normal(dim).astype(np.float32).tolist()
    n, done, t_total = 256, 0, 0.0
    scores = []
    while t_total < target_seconds and done < rows:
        X = rng.standard_normal((n, dim)).astype(np.float32).tolist()
        t0 = time.perf_counter()
        scores.extend((O.cosine_pair(q, x), done + i) for i, x in enumerate(X))
        t_total += time.perf_counter() - t0
        done += n
    t0 = time.perf_counter()
    sorted(scores, key=lambda t: t[0], reverse=True)[:k]
    t_total += time.perf_counter() - t0
    return {"value": 1.0 / (t_total * rows / done), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 query x {done} rows x {dim}, pure-Python _cosine restatement + stable sort, {t_total:.2f} s; scaled by {rows}/{done} rows"}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    qps, cores, sample, t = cpu_reference_qps(a.rows, a.dim, a.batch, a.k, target_seconds=max(20.0, a.cpu_seconds),
                                              steps=a.steps, warmup=a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": shared_config(a),
        "detail": {"extrapolated": True,
                   "how": "the CPU arm times a bounded sample (<= 256 queries x a row subset) and scales queries/s linearly in rows to the "
                          "full corpus; see cpu_baseline.sample"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python (no compiled path); its numpy/BLAS oracle port is timed, all host threads",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# recall leg: the oracle as the CHECKER of the results the timed regions produced (outside any timed region)
# ------------------------------------------------------------------------------------------------
def oracle_lists_for_shard(X, lo, Qsel_f32, k, dtype_name, block_rows=1 << 18, extra=64):
    """Exact fp64 top-(k+extra) of the selected queries over THIS rank's corpus shard, streamed back from HBM in row
    blocks (global ids = lo + local row)."""
    import numpy as np
    from oracle import retrieval_oracle as O

    def blocks():
        n = int(X.shape[0])
        for r0 in range(0, n, block_rows):
            xb = X[r0:min(n, r0 + block_rows)]
            if dtype_name == "bf16":
                import torch
                yield lo + r0, O.bf16_bits_to_f32(xb.view(torch.int16).cpu().numpy().view(np.uint16))
            else:
                yield lo + r0, xb.cpu().numpy().astype(np.float32)
    return O.topk_blocks_exact(Qsel_f32, blocks(), k, metric="cos", extra=extra)


def judge_results(ref_s, ref_i, results, k, rel_tol=1e-3):
    """results: {name: (scores [nq, k], ids [nq, k])} -> per-result worst case over the checked queries."""
    from oracle import retrieval_oracle as O
    out = {}
    for name, (gs, gi) in results.items():
        rs = [O.check_topk_against_list(gs[j], gi[j], ref_s[j], ref_i[j], k, rel_tol) for j in range(ref_s.shape[0])]
        out[name] = {"recall": min(r["recall"] for r in rs), "recall_strict": min(r["recall_strict"] for r in rs),
                     "max_rel_score_err": max(r["max_rel_score_err"] for r in rs), "ok": all(r["ok"] for r in rs)}
    return out



# ------------------------------------------------------------------------------------------------
# secondary legs (after the contract timing, reported under "legs"): the other north-star targets on the SAME run
# ------------------------------------------------------------------------------------------------
def leg_small_batch(torch, index, world, Q, k, steps, n_local, dim, peaks, _native):
    """Latency of ONE small batch end to end on the device (search + exchange + merge at N > 1), and the HBM roofline of
    its tile kernel (one query tile: the corpus shard is streamed once)."""
    B = int(Q.shape[1])
    for i in range(3):
        index.search(Q[i % Q.shape[0]], k)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        index.search(Q[i % Q.shape[0]], k)
    e1.record()
    torch.cuda.synchronize()
    kms, kn = _native.profile_read()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    kms /= steps
    gbs = n_local * dim * 2 / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    peak = peaks.get("hbm_gbs", 6650.0)
    # the same step replayed as ONE CUDA graph launch (search + exchange + merge): what a latency-bound caller would use
    graph_ms, graph_err = None, None
    try:
        from a_modular_rag_framework_b200.index import CapturedSearch
        cap = CapturedSearch(index, Q[0], k)
        for i in range(3):
            cap(Q[i % Q.shape[0]])
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            cap(Q[i % Q.shape[0]])
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / steps
    except Exception as e:  # noqa: BLE001
        graph_err = repr(e)
    return {"batch": B, "k": k, "ms_per_batch": ms, "queries_per_s": B / (ms * 1e-3), "tile_kernel_ms": kms, "launches_per_batch": kn / steps,
            "cuda_graph_ms_per_batch": graph_ms, "cuda_graph_queries_per_s": (B / (graph_ms * 1e-3)) if graph_ms else None,
            "cuda_graph_error": graph_err,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "algorithmic": "N_local*D*2 bytes per batch per GPU (corpus shard streamed once), tile kernel time only"}}


def gen_shard(torch, lo, hi, dim, tdt, dev, blk=1 << 20):
    """Rows [lo, hi) of the synthetic unit-norm corpus on `dev`: Gaussian rows seeded per GLOBAL 2^20-row-aligned piece of
    the range, so any process that generates rows [lo, hi) with the same (lo, blk) gets the same bits."""
    n_local = hi - lo
    X = torch.empty((max(n_local, 1), dim), dtype=tdt, device=dev)
    for r0 in range(lo, hi, blk):
        r1 = min(hi, r0 + blk)
        g = torch.Generator(device=dev).manual_seed(1_000_003 + r0)
        t = torch.randn((r1 - r0, dim), generator=g, device=dev, dtype=torch.float32)
        t /= t.norm(dim=1, keepdim=True)
        X[r0 - lo: r1 - lo] = t.to(tdt)
        del t
    if n_local == 0:
        X.zero_()
    return X


def leg_plugin_single_process(torch, a, world, tdt, Qh, ref_ids, steps, warmup):
    """N > 1, rank 0 only, while the other ranks wait: ONE process drives all N GPUs through the plug-in's engine
    (backend.CudaEngine, what B200RetrievalBackend(devices=[0..N-1]) builds): float32 host queries in, host results out, the
    corpus row-sharded over the devices, exchange + merge inside libb200r (b200r_search_sharded_host_all_f32)."""
    import numpy as np
    from a_modular_rag_framework_b200.backend import CudaEngine
    from a_modular_rag_framework_b200.sharded import shard_rows
    shards = []
    for g in range(world):
        lo, hi = shard_rows(a.rows, world, g)
        shards.append(gen_shard(torch, lo, hi, a.dim, tdt, torch.device("cuda", g)))
    eng = CudaEngine.from_shards(shards, a.rows, "cos", max_batch=a.batch, max_k=a.k)
    Qf = [Qh[i].float().numpy() for i in range(Qh.shape[0])]
    for i in range(min(warmup, 2)):
        eng.dense_search(Qf[i], a.k)
    t0 = time.perf_counter()
    for i in range(steps):
        s, ids = eng.dense_search(Qf[warmup + i], a.k)
    dt = time.perf_counter() - t0
    same = bool((torch.from_numpy(ids) == ref_ids.cpu()).all().item())
    # one query at a time, the way the reference's orchestrator calls its retriever (app/orchestrator/nodes.py:110-115)
    q1 = [np.ascontiguousarray(Qf[i % len(Qf)][:1]) for i in range(8)]
    for i in range(5):
        eng.dense_search(q1[i % 8], a.k)
    t1 = time.perf_counter()
    n1 = 100
    for i in range(n1):
        eng.dense_search(q1[i % 8], a.k)
    b1_ms = (time.perf_counter() - t1) / n1 * 1e3
    del eng, shards
    torch.cuda.empty_cache()
    return {"value": steps * a.batch / dt, "unit": UNIT, "ms_per_step": dt / steps * 1e3, "devices": world,
            "h2d_bytes_per_step": a.batch * a.dim * 4 * world, "d2h_bytes_per_step": a.batch * a.k * 12,
            "identical_to_the_torchrun_result": same, "batch1_ms_per_query_host_to_host": b1_ms,
            "note": "one host thread, N GPUs: float32 host queries, async H2D to every device, rounded to the corpus dtype on the device, "
                    "sharded search + exchange + merge (libb200r), D2H from device 0; not pipelined across batches"}


def build_graph(torch, V, E, dev, seed=5):
    """Synthetic passage graph, identical on every rank: E random directed edges, symmetrised, neighbours sorted."""
    g = torch.Generator(device=dev).manual_seed(seed)
    src = torch.randint(0, V, (E,), generator=g, device=dev)
    dst = torch.randint(0, V, (E,), generator=g, device=dev)
    a, b = torch.cat([src, dst]), torch.cat([dst, src])
    del src, dst
    order = torch.argsort(a * V + b)
    a, b = a[order], b[order]
    del order
    row_ptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(a, minlength=V), 0)
    return row_ptr, b.to(torch.int32).contiguous()


def leg_graph(torch, dist, a, X, rank, world, dev, Qd, peaks, _native, steps=20, warmup=5):
    """BASELINE.json configs[4]: dense top-20 seeds -> 2-hop CSR expansion over a 5M-node / 50M-edge graph -> re-score ->
    fused top-10.  The graph corpus is the first rows of every rank's shard (row-sharded like the dense corpus); at
    N > 1 the stage runs over the library's exchange (queries sharded by home rank, candidates by owner rank)."""
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore
    from a_modular_rag_framework_b200.comm import ShardComm
    from a_modular_rag_framework_b200.index import expand_rescore_exchange
    from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, shard_rows
    V = min(a.graph_nodes, a.rows // 2)
    E = int(a.graph_edges * (V / a.graph_nodes))
    B, S, k, window = a.batch, 20, 10, 2
    glo, ghi = shard_rows(V, world, rank)
    Xg = X[: max(ghi - glo, 1)]
    local = DenseIndex(Xg, metric="cos", id_base=glo)
    row_ptr, col = build_graph(torch, V, E, dev)
    graph = PassageGraph(row_ptr, col)
    nb = min(4, Qd.shape[0])
    if world > 1:
        comm = ShardComm.from_process_group(dev.index, B, S, max_graph_k=k)
        sidx = ShardedDenseIndex(local, exchange="abi", comm=comm)
        seeds = [sidx.search(Qd[i], S)[1].contiguous() for i in range(nb)]
        run = lambda i: expand_rescore_exchange(local, graph, comm, Qd[i % nb], seeds[i % nb], window, 0.4, 0.2, k, V)   # noqa: E731
    else:
        seeds = [local.search(Qd[i], S)[1].contiguous() for i in range(nb)]
        run = lambda i: expand_rescore(local, graph, Qd[i % nb], seeds[i % nb], window, 0.4, 0.2, k, check_status=False)   # noqa: E731
    for i in range(warmup):
        out = run(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = run(i)
    e1.record()
    torch.cuda.synchronize()
    gms, gn = _native.profile_read()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # per-launch-site timeline of the stage (events around every launch, 5 extra batches)
    timeline = None
    try:
        _native.profile_timeline(True)
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        timeline = {site: {"ms_per_batch": v[0] / 5, "launches_per_batch": v[1] / 5} for site, v in _native.profile_dump().items()}
    except Exception as e:  # noqa: BLE001
        timeline = {"error": repr(e)}
    finally:
        _native.profile_timeline(False)
    ncand = float(out["ncand"].float().mean().item())
    gather_bytes = ncand * B * a.dim * 2 / world               # per GPU
    peak = peaks.get("hbm_gbs", 6650.0)
    gms = gms / max(gn, 1)
    leg = {"workload": f"{V} nodes / {2 * E} adjacency entries, {a.dim}-d {a.dtype} rows, batch {B}, {S} dense seeds, window {window}, top-{k}",
           "ms_per_batch": ms, "queries_per_s": B / (ms * 1e-3), "avg_candidates": ncand,
           "overflowed_queries": int((out["status"] != 0).sum().item()),
           "roofline": {"bound": "hbm", "achieved": gather_bytes / (gms * 1e-3) / 1e9 if gms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                        "frac": (gather_bytes / (gms * 1e-3) / 1e9 / peak) if gms > 0 else 0.0, "kernel": "rescore_gather_kernel",
                        "kernel_ms_per_batch": gms,
                        "algorithmic": "candidates*D*2 bytes per batch per GPU (random 1.5 KB rows), gather kernel time only"},
           "stage_hbm_frac": gather_bytes / (ms * 1e-3) / 1e9 / peak, "timeline": timeline}
    # ---- parity of the last batch against the oracle (2 queries): BFS on the host CSR, rows fetched from their owners ----
    try:
        leg["verify"] = verify_graph_leg(torch, dist, out, Qd[(steps - 1) % nb], seeds[(steps - 1) % nb], row_ptr, col, Xg, glo, ghi,
                                         V, rank, world, window, k)
    except Exception as e:  # noqa: BLE001
        leg["verify"] = {"ok": False, "error": repr(e)}
    return leg


def verify_graph_leg(torch, dist, out, Q, seeds, row_ptr, col, Xg, glo, ghi, V, rank, world, window, k, sel=(0, 1)):
    import numpy as np
    from oracle import retrieval_oracle as O
    sel = [q for q in sel if q < Q.shape[0]]
    rp = row_ptr.cpu().numpy() if rank == 0 else None
    ci = col.cpu().numpy() if rank == 0 else None
    t_cpu = time.perf_counter()
    dists = [O.bfs_expand(rp, ci, seeds[q].cpu().numpy(), window) for q in sel] if rank == 0 else None
    t_cpu = time.perf_counter() - t_cpu            # the oracle's BFS (expand_qmatch_neighbors restated), one host core
    if world > 1:
        box = [dists]
        dist.broadcast_object_list(box, src=0)
        dists = box[0]
    # every rank contributes the (dequantised) rows it owns
    mine = {}
    for d in dists:
        for node in d:
            if glo <= node < ghi and node not in mine:
                mine[node] = Xg[node - glo].float().cpu().numpy()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    else:
        parts = [mine]
    if rank != 0:
        return None
    rows = {}
    for p_ in parts:
        rows.update(p_)
    res = {"queries": len(sel), "ok": True, "max_rel_cos_err": 0.0, "ids_equal": True}
    for j, q in enumerate(sel):
        d = dists[j]
        cand = sorted(d)
        Xc = np.stack([rows[c] for c in cand]).astype(np.float64)
        t1 = time.perf_counter()
        cos = O.score_matrix(Q[q].float().cpu().numpy()[None].astype(np.float64), Xc)[0]
        dense = dict(zip(cand, cos.tolist()))
        full = O.fuse_and_rank(dense, {c: O.decay(d[c]) for c in cand}, 0.4, 0.2, len(cand))
        t_cpu += time.perf_counter() - t1          # + re-score of the candidates (numpy) + min-max fusion + ranking
        ref_of = {t[0]: t[1] for t in full}
        scale = max(1e-30, float(np.max(np.abs(cos))))
        rng_ = float(cos.max() - cos.min())
        tol_f = (0.4 * 2 * 1e-3 * scale / rng_ if rng_ > 0 else 0.0) + 1e-6
        ids = out["id"][q].cpu().numpy()[:k]
        sc = out["score"][q].cpu().numpy()[:k]
        cv = out["cos"][q].cpu().numpy()[:k]
        hop = out["hop"][q].cpu().numpy()[:k]
        if int(out["ncand"][q]) != len(cand):
            res["ok"] = False
        for i, nid in enumerate(ids.tolist()):
            if nid not in d or hop[i] != d[nid] or abs(cv[i] - dense[nid]) > 1e-3 * scale or abs(sc[i] - ref_of[nid]) > tol_f:
                res["ok"] = False
                continue
            res["max_rel_cos_err"] = float(max(res["max_rel_cos_err"], abs(float(cv[i]) - dense[nid]) / scale))
            if nid != full[i][0]:
                res["ids_equal"] = False
                if abs(ref_of[nid] - full[i][1]) > tol_f:       # not a near-tie
                    res["ok"] = False
    if sel and t_cpu > 0:
        res["cpu_reference"] = {"value": len(sel) / t_cpu, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"{len(sel)} queries of the batch: BFS + decay in Python over the host CSR, numpy fp64 re-score of "
                                          f"the candidates, min-max fusion, ranking (oracle port of graph_utils.py:58-129 + "
                                          f"retrieval_backend.py:296-372; rows already gathered)"}
    return res


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
        0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} must be launched with torchrun --nproc-per-node {a.gpus}")
        a.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from a_modular_rag_framework_b200 import DenseIndex, _native
    from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, shard_rows

    # ---- synthetic corpus shard, generated on the device in row blocks (seeded per block) -----------
    lo, hi = shard_rows(a.rows, world, rank)
    n_local = hi - lo
    tdt = torch.bfloat16 if a.dtype == "bf16" else torch.float16
    X = gen_shard(torch, lo, hi, a.dim, tdt, dev)
    n_batches = a.warmup + a.steps
    gq = torch.Generator(device="cpu").manual_seed(4242)
    Qh = torch.randn((n_batches, a.batch, a.dim), generator=gq, dtype=torch.float32)
    Qh /= Qh.norm(dim=2, keepdim=True)
    Qh = Qh.to(tdt).pin_memory()        # same queries on every rank (replicated)
    Qd = Qh.to(dev)
    torch.cuda.synchronize()

    local = DenseIndex(X[:max(n_local, 1)] if n_local else X, metric="cos", id_base=lo)
    index = ShardedDenseIndex(local, exchange=a.exchange) if world > 1 else local
    if world > 1 and a.exchange == "abi":
        # the exchange windows are mapped across the ranks with CUDA IPC; every rank learns the outcome of the bring-up
        try:
            index.ensure_comm(a.batch, a.k)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"[bench] rank {rank}: C-ABI exchange unavailable ({e!r}); using the NCCL all-gather\n")
            a.exchange = "nccl"
            index = ShardedDenseIndex(local, exchange="nccl")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------
    for i in range(a.warmup):
        r = index.search(Qd[i], a.k, path=a.path, overlap=True) if world > 1 else index.search(Qd[i], a.k, path=a.path)
    if world > 1:
        r.get()
    sampler = ClockSampler(local_rank)        # NVML init takes a variable time per process: do it BEFORE the barrier
    _native.profile_enable(True)
    launches0 = _native.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    pending = None
    for i in range(a.steps):
        if world > 1:
            # throughput mode: this batch's all-gather + merge overlap the next batch's local search
            pending = index.search(Qd[a.warmup + i], a.k, path=a.path, overlap=True)
        else:
            out = index.search(Qd[a.warmup + i], a.k, path=a.path)
    if pending is not None:
        out = pending.get()       # the timed region ends when the LAST batch's merged result is ready
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = _native.launch_count() - launches0
    kernel_ms, kernel_n = _native.profile_read()
    _native.profile_enable(False)
    ms = ev0.elapsed_time(ev1)
    per_rank = [{"rank": rank, "ms_per_step": ms / a.steps, "kernel_ms_per_step": kernel_ms / a.steps}]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
    value = a.steps * a.batch / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        hs = torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory()
        hi_ = torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory()
        if world == 1:
            def e2e_step(i):
                local.search_host(Qh[i], a.k, path=a.path, out=(hs, hi_))

            def e2e_finish():
                pass
        else:
            # software-pipelined by one batch: while batch i is searched, batch i-1's exchange/merge finishes on the side
            # stream and its results are copied to the host; every step still does its own H2D and D2H
            copy_stream = torch.cuda.Stream(device=dev)
            inflight = []

            def drain(n_keep):
                while len(inflight) > n_keep:
                    pend = inflight.pop(0)
                    with torch.cuda.stream(copy_stream):
                        s, ids = pend.get()
                        hs.copy_(s, non_blocking=True)
                        hi_.copy_(ids, non_blocking=True)
                        done = torch.cuda.Event()
                        done.record(copy_stream)
                    done.synchronize()          # the host has this batch's results

            def e2e_step(i):
                q = Qh[i].to(dev, non_blocking=True)
                inflight.append(index.search(q, a.k, path=a.path, overlap=True))
                drain(1)

            def e2e_finish():
                drain(0)
        for i in range(min(a.warmup, 2)):
            e2e_step(i)
        e2e_finish()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            e2e_step(a.warmup + i)
        e2e_finish()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": a.steps * a.batch / dt, "unit": UNIT, "h2d_bytes_per_step": a.batch * a.dim * 2,
               "d2h_bytes_per_step": a.batch * a.k * 12, "ms_per_step": dt / a.steps * 1e3,
               "note": "pinned host queries -> H2D -> search -> D2H results to pinned host, per step; corpus index resident"
                       + ("" if world == 1 else "; batches pipelined by one (exchange/merge/D2H of batch i-1 overlap the search of batch i)")}

    # ---- recall leg (outside the timed regions): oracle top-k of a few queries of the LAST timed batch over the whole
    # (sharded) corpus, streamed back from HBM; checks the device-resident result and the e2e host result ------------
    verify = None
    if a.verify_queries > 0:
        import numpy as np
        t0 = time.perf_counter()
        nq = min(a.verify_queries, a.batch)
        sel = np.unique(np.linspace(0, a.batch - 1, nq).astype(np.int64))
        last = a.warmup + a.steps - 1
        Qsel = Qh[last][torch.from_numpy(sel)].float().numpy()
        loc_s, loc_i = oracle_lists_for_shard(X[:n_local], lo, Qsel, a.k, a.dtype)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (loc_s, loc_i))
        else:
            parts = [(loc_s, loc_i)]
        if rank == 0:
            from oracle import retrieval_oracle as O
            ref_s, ref_i = O.merge_topk(np.stack([p_[0] for p_ in parts], axis=1), np.stack([p_[1] for p_ in parts], axis=1),
                                        loc_s.shape[1])
            results = {"device": (out[0][sel].cpu().numpy(), out[1][sel].cpu().numpy())}
            if e2e is not None:
                results["e2e"] = (hs.numpy()[sel], hi_.numpy()[sel])
            verify = judge_results(ref_s, ref_i, results, a.k)
            verify["queries"] = int(len(sel))
            verify["seconds"] = time.perf_counter() - t0
            verify["how"] = ("fp64 cosine + stable (score desc, id asc) order of the oracle over all rows of every shard (f32 BLAS "
                             "prefilter per 262144-row block, fp64 re-score), merged over ranks; tie-aware rule, 1e-3 relative")

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    used_path = a.path if a.path != "auto" else "gemm"
    # the dominant kernel is launched once per phase of a step: sum its launches over the step
    kernel_avg_ms = kernel_ms / max(a.steps, 1)
    launches_per_step = kernel_n / max(a.steps, 1)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            traffic, traffic_note = tj.get(f"{used_path}_bytes_per_launch"), tj.get(f"{used_path}_note")
    except Exception:
        pass
    hbm_bound_tc = used_path == "gemm" and a.batch <= 128     # one query tile: the tcgen05 kernels stream the corpus once
    if used_path == "gemm" and not hbm_bound_tc:
        flops = 2.0 * a.batch * n_local * a.dim
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = flops / (kernel_avg_ms * 1e-3) / 1e12 if kernel_avg_ms > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "kernel": "gemm_topk_kernel", "kernel_ms_per_step": kernel_avg_ms, "launches_per_step": launches_per_step,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                    "algorithmic": "2*B*N_local*D flop per step, summed over the phase launches of gemm_topk_kernel"}
    elif hbm_bound_tc:
        nbytes = float(n_local) * a.dim * 2
        peak = peaks.get("hbm_gbs", 6650.0)
        achieved = nbytes / (kernel_avg_ms * 1e-3) / 1e9 if kernel_avg_ms > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": "gemm_topk_kernel<1>",
                    "kernel_ms_per_step": kernel_avg_ms, "launches_per_step": launches_per_step,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                    "algorithmic": "N_local*D*2 bytes per step (corpus streamed once), summed over the phase launches"}
    else:
        nbytes = float(n_local) * a.dim * 2 * ((a.batch + 3) // 4)
        peak = peaks.get("hbm_gbs", 6650.0)
        achieved = nbytes / (kernel_avg_ms * 1e-3) / 1e9 if kernel_avg_ms > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "scan_topk_kernel", "kernel_ms_per_step": kernel_avg_ms, "launches_per_step": launches_per_step,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                    "algorithmic": "N_local*D*2 bytes per corpus pass (one pass per <=4 queries), summed over the step"}

    # ---- per-launch-site timeline of the dense step (events around every launch; 5 extra steps, outside the timed region) ----
    timeline = None
    try:
        _native.profile_timeline(True)
        for i in range(5):
            if world > 1:
                index.search(Qd[a.warmup + i % a.steps], a.k, path=a.path, overlap=True)
            else:
                index.search(Qd[a.warmup + i % a.steps], a.k, path=a.path)
        if world > 1 and getattr(index, "_outstanding", None) is not None:
            index._outstanding.resolve()
            index._outstanding = None
        torch.cuda.synchronize()
        timeline = {site: {"ms_per_step": v[0] / 5, "launches_per_step": v[1] / 5} for site, v in _native.profile_dump().items()}
    except Exception as e:  # noqa: BLE001
        timeline = {"error": repr(e)}
    finally:
        _native.profile_timeline(False)

    # ---- secondary legs: batch-1 / batch-16 latency on the same shards, graph stage (configs[4]) --------------------------
    legs = None
    if not a.no_legs:
        legs = {}
        if world > 1 and a.exchange == "abi" and getattr(index, "_outstanding", None) is not None:
            index._outstanding.resolve()
            index._outstanding = None
        gs = torch.Generator(device="cpu").manual_seed(777)
        for Bs in (1, 16):
            Qs = torch.randn((8, Bs, a.dim), generator=gs, dtype=torch.float32)
            Qs /= Qs.norm(dim=2, keepdim=True)
            legs[f"batch{Bs}"] = leg_small_batch(torch, index, world, Qs.to(tdt).to(dev), a.k, 50, n_local, a.dim, peaks, _native)
        try:
            legs["graph"] = leg_graph(torch, dist, a, X, rank, world, dev, Qd, peaks, _native)
        except Exception as e:  # noqa: BLE001
            legs["graph"] = {"error": repr(e)}
        if world > 1:
            # the other ranks wait on the HOST (gloo): an NCCL barrier would park a spinning kernel on their GPUs, which rank 0's
            # kernels would have to time-slice with
            torch.cuda.synchronize()
            cpu_group = dist.new_group(backend="gloo")
            dist.barrier(group=cpu_group)
            if rank == 0 and not a.no_plugin_leg:
                try:
                    legs["plugin_single_process"] = leg_plugin_single_process(torch, a, world, tdt, Qh, out[1], a.steps, a.warmup)
                except Exception as e:  # noqa: BLE001
                    legs["plugin_single_process"] = {"error": repr(e)}
                torch.cuda.set_device(local_rank)
            dist.barrier(group=cpu_group)

    # ---- CPU baseline on rank 0, N=1 only ---------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del out
        qps, cores, sample, _ = cpu_reference_qps(a.rows, a.dim, a.batch, a.k, a.cpu_seconds)
        cpu_baseline = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                        "reference_style_python": cpu_python_style_qps(a.rows, a.dim, a.k)}

    failed = False
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": shared_config(a),
            "detail": {"path": used_path, "sharding": f"rows/{world}",
                       "exchange": "none" if world == 1 else (
                           "libb200r exchange (b200r_comm_*): the final select kernel stores each rank's packed top-k into every rank's window over NVLink (CUDA IPC mapped), device-side epoch flags, one merge kernel queued behind the next batch's search; no NCCL in the data path"
                           if a.exchange == "abi" else
                           "packed NCCL all-gather + device merge on a side stream, overlapped with the next batch's search")},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "per_rank": per_rank, "legs": legs, "timeline": timeline,
        }
        if verify is not None:
            line["recall"] = min(v["recall"] for v in verify.values() if isinstance(v, dict))
            line["max_rel_score_err"] = max(v["max_rel_score_err"] for v in verify.values() if isinstance(v, dict))
            line["verify"] = verify
            failed = not all(v["ok"] for v in verify.values() if isinstance(v, dict))
        gv = ((legs or {}).get("graph") or {}).get("verify")
        if isinstance(gv, dict) and not gv.get("ok", False):
            failed = True
        print(json.dumps(line, default=float), flush=True)
    if world > 1:
        flag = torch.tensor([1 if (rank == 0 and failed) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        failed = bool(int(flag.item()))
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.stderr.write("[bench] recall check FAILED: the timed results differ from the oracle\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
