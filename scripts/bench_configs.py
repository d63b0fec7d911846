"""Secondary measurements (not the driver's contract line): the other BASELINE.json configs on ONE GPU.
Each line reports queries/s and the roofline fraction of the dominant kernel (CUDA events via the
library's profiling hooks).  Usage: python scripts/bench_configs.py [names...]"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore, _native

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def sm_clock_mhz():
    """SM clock right now (called while the timed kernels are still queued): several kernels here are bound by an SM-side
    pipeline, so a power-capped box measures slower -- the line says which clock it ran at."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        return int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception:  # noqa: BLE001
        return None


def corpus(N, D, dtype, seed=1):
    X = torch.empty((N, D), dtype=dtype, device=dev)
    blk = 1 << 20
    for r0 in range(0, N, blk):
        r1 = min(N, r0 + blk)
        g = torch.Generator(device=dev).manual_seed(seed + r0)
        t = torch.randn((r1 - r0, D), generator=g, device=dev)
        t /= t.norm(dim=1, keepdim=True)
        X[r0:r1] = t.to(dtype)
    return X


def queries(n, B, D, dtype, seed=99):
    g = torch.Generator(device=dev).manual_seed(seed)
    Q = torch.randn((n, B, D), generator=g, device=dev)
    Q /= Q.norm(dim=2, keepdim=True)
    return Q.to(dtype)


def run_search(name, N, D, B, k, dtype, path, steps=10, warmup=3, check=True):
    X = corpus(N, D, dtype)
    Q = queries(steps + warmup, B, D, dtype)
    ix = DenseIndex(X, metric="cos")
    for i in range(warmup):
        out = ix.search(Q[i], k, path=path)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = ix.search(Q[warmup + i], k, path=path)
    e1.record()
    mhz = sm_clock_mhz()
    torch.cuda.synchronize()
    kms, kn = _native.profile_read()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    kms /= steps
    esz = X.element_size()
    used = path if path != "auto" else ("gemm" if dtype in (torch.bfloat16, torch.float16) else "scan")
    line = {"config": name, "rows": N, "dim": D, "batch": B, "k": k, "dtype": str(dtype).split(".")[-1], "path": used,
            "ms_per_batch": ms, "queries_per_s": B / (ms * 1e-3), "kernel_ms_per_batch": kms, "kernel_launches_per_batch": kn / steps, "sm_mhz": mhz}
    # corpus passes per batch: the tile kernel streams the corpus once; the bulk-fed fp32 scan serves 8 queries per pass
    # (scan_bulk.cu), the register-fed scan 4 (scan_topk.cu)
    per_pass = 8 if dtype == torch.float32 else 4
    passes = 1 if used == "gemm" else (B + per_pass - 1) // per_pass
    if used == "gemm" and dtype == torch.float32:
        passes = (B + 255) // 256        # rows_f32.cu: 16 queries per group, the groups of a 256-query slice ganged over L2
    gbs = N * D * esz * passes / (kms * 1e-3) / 1e9
    tfs = 2.0 * B * N * D / (kms * 1e-3) / 1e12
    line["hbm_GBps"] = gbs
    line["hbm_frac_of_measured"] = gbs / PEAKS["hbm_gbs"]
    line["tensor_TFLOPs"] = tfs
    line["tensor_frac_of_measured_sustained"] = tfs / PEAKS["bf16_tflops_sustained"]
    if check:
        from oracle import retrieval_oracle as O
        s, ids = out
        nq = min(B, 2)
        Qn = Q[warmup + steps - 1][:nq].float().cpu().numpy()
        best_s = np.full((nq, 0), 0.0); best_i = np.zeros((nq, 0), dtype=np.int64)
        for r0 in range(0, N, 1 << 20):
            Xn = X[r0:r0 + (1 << 20)].float().cpu().numpy()
            ss, ii = O.topk(Qn, Xn, k, dtype=np.float32, id_base=r0)
            best_s = np.concatenate([best_s, ss], 1); best_i = np.concatenate([best_i, ii], 1)
        ms_, mi_ = O.merge_topk(best_s[:, None, :], best_i[:, None, :], k)
        got = ids[:nq].cpu().numpy()
        inter = np.mean([len(set(got[b]) & set(mi_[b])) / k for b in range(nq)])
        line["recall_vs_oracle_f32"] = float(inter)
        line["max_abs_score_err"] = float(np.max(np.abs(np.sort(s[:nq].cpu().numpy(), 1) - np.sort(ms_, 1))))
    print(json.dumps(line), flush=True)
    del ix, X, Q
    torch.cuda.empty_cache()
    return line


def run_search_i8(name, N, D, B, k, steps=10, warmup=3, path="auto", check=True):
    Xf = corpus(N, D, torch.float16)
    X = torch.empty((N, D), dtype=torch.int8, device=dev)
    blk = 1 << 20
    for r0 in range(0, N, blk):
        X[r0:r0 + blk] = torch.clamp(torch.round(Xf[r0:r0 + blk].float() * 127.0 / Xf[r0:r0 + blk].float().abs().amax(dim=1, keepdim=True)), -127, 127).to(torch.int8)
    del Xf
    Q = queries(steps + warmup, B, D, torch.float32)
    ix = DenseIndex(X, metric="cos")
    for i in range(warmup):
        ix.search(Q[i], k, path=path)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ix.search(Q[warmup + i], k, path=path)
    e1.record()
    torch.cuda.synchronize()
    kms, kn = _native.profile_read()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    kms /= steps
    tensor = path != "scan"
    # tcgen05 kind::i8: the (up to 4) 64-query groups of a 256-query slice share one launch and the corpus tiles through L2,
    # so HBM sees the corpus once per slice; the MMA work is three signed-byte limbs per query
    passes = (B + 255) // 256 if tensor else (B + 3) // 4
    gbs = N * D * passes / (kms * 1e-3) / 1e9
    tops = 3 * 2.0 * B * N * D / (kms * 1e-3) / 1e12
    line = {"config": name, "rows": N, "dim": D, "batch": B, "k": k, "dtype": "int8", "path": "tensor_i8" if tensor else "scan", "ms_per_batch": ms,
            "kernel_launches_per_batch": kn / steps, "queries_per_s": B / (ms * 1e-3), "kernel_ms_per_batch": kms, "hbm_GBps": gbs,
            "hbm_frac_of_measured": gbs / PEAKS["hbm_gbs"]}
    if tensor:
        line["int8_limb_TOPs"] = tops
        line["frac_of_2x_bf16_sustained"] = tops / (2.0 * PEAKS["bf16_tflops_sustained"])
    if check:
        # last batch, a few queries spread over the groups: fp32 scores of the dequantised rows on the device (torch, test
        # infrastructure), ranked by the oracle's rule
        from oracle import retrieval_oracle as O
        s, ids = ix.search(Q[warmup + steps - 1], k, path=path)
        sel = sorted(set(int(x) for x in np.linspace(0, B - 1, num=min(B, 4))))
        Qs = Q[warmup + steps - 1][sel].float()
        Qs = Qs / Qs.norm(dim=1, keepdim=True)
        sc = torch.empty((len(sel), N), dtype=torch.float32, device=dev)
        for r0 in range(0, N, blk):
            Xb = X[r0:r0 + blk].float()
            sc[:, r0:r0 + blk] = (Qs @ Xb.T) / Xb.norm(dim=1)[None, :]
        ws_, wi_ = O.topk_from_scores(sc.cpu().numpy().astype(np.float64), k)
        got_i, got_s = ids[sel].cpu().numpy(), s[sel].cpu().numpy()
        line["recall_vs_f32"] = float(np.mean([len(set(got_i[b]) & set(wi_[b])) / k for b in range(len(sel))]))
        line["max_abs_score_err"] = float(np.max(np.abs(np.sort(got_s, 1) - np.sort(ws_, 1))))
    print(json.dumps(line), flush=True)
    del ix, X, Q
    torch.cuda.empty_cache()


def run_graph(name, V, E, D, B, S, k, steps=40, warmup=8, window=2, random_seeds=False, nbatch=4):
    X = corpus(V, D, torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(5)
    src = torch.randint(0, V, (E,), generator=g, device=dev)
    dst = torch.randint(0, V, (E,), generator=g, device=dev)
    a = torch.cat([src, dst]); b = torch.cat([dst, src])
    order = torch.argsort(a * V + b)
    a, b = a[order], b[order]
    row_ptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(a, minlength=V), 0)
    col = b.to(torch.int32).contiguous()
    del src, dst, a, b, order
    graph = PassageGraph(row_ptr, col)
    ix = DenseIndex(X, metric="cos")
    Q = queries(nbatch, B, D, torch.bfloat16)
    if random_seeds:   # window=0 + S random seeds = the re-score gather alone (no BFS)
        seeds = [torch.randint(0, V, (B, S), generator=g, device=dev) for i in range(nbatch)]
    else:              # seeds = the dense top-S of each query (the north-star pipeline)
        seeds = [ix.search(Q[i], S)[1] for i in range(nbatch)]
    for i in range(warmup):
        out = expand_rescore(ix, graph, Q[i % nbatch], seeds[i % nbatch], window, 0.4, 0.2, k)
    torch.cuda.synchronize()
    _native.profile_enable(True)        # brackets the gather kernel (K3b)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        out = expand_rescore(ix, graph, Q[i % nbatch], seeds[i % nbatch], window, 0.4, 0.2, k, check_status=False)
        ev[i + 1].record()
    torch.cuda.synchronize()
    gms, gn = _native.profile_read()
    _native.profile_enable(False)
    gms /= max(gn, 1)
    # per-launch-site timeline of a few more batches (events around every launch: not part of the timed loop above)
    _native.profile_timeline(True)
    for i in range(4):
        expand_rescore(ix, graph, Q[i % nbatch], seeds[i % nbatch], window, 0.4, 0.2, k, check_status=False)
    torch.cuda.synchronize()
    timeline = {name: round(v[0] / 4 * 1000, 1) for name, v in _native.profile_dump().items()}
    _native.profile_timeline(False)
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    ms = ev[0].elapsed_time(ev[steps]) / steps
    ncand = out["ncand"].float().mean().item()
    gather_bytes = ncand * B * D * 2
    line = {"config": name, "nodes": V, "edges_undirected": 2 * E, "dim": D, "batch": B, "seeds": S, "k": k, "window": window,
            "steps": steps, "ms_per_batch": ms, "ms_median": per[len(per) // 2], "ms_min": per[0], "queries_per_s": B / (ms * 1e-3), "avg_candidates": ncand,
            "gather_GBps": gather_bytes / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": gather_bytes / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"],
            "gather_kernel_ms": gms, "gather_kernel_GBps": gather_bytes / (gms * 1e-3) / 1e9 if gms > 0 else None,
            "overflowed_queries": int((out["status"] != 0).sum().item()), "timeline_us_per_batch": timeline}
    print(json.dumps(line), flush=True)


CONFIGS = {
    "cfg1_1Mx768_B1024": lambda: run_search("cfg1_1Mx768_B1024", 1_000_000, 768, 1024, 100, torch.bfloat16, "auto"),
    "cfg2_10Mx768_B1_scan": lambda: run_search("cfg2_10Mx768_B1_scan", 10_000_000, 768, 1, 100, torch.bfloat16, "scan", steps=20),
    "cfg2_10Mx768_B4_scan": lambda: run_search("cfg2_10Mx768_B4_scan", 10_000_000, 768, 4, 100, torch.bfloat16, "scan", steps=20),
    "cfg2_10Mx768_B16_gemm": lambda: run_search("cfg2_10Mx768_B16_gemm", 10_000_000, 768, 16, 100, torch.bfloat16, "gemm", steps=20),
    "cfg2_10Mx768_B4_gemm": lambda: run_search("cfg2_10Mx768_B4_gemm", 10_000_000, 768, 4, 100, torch.bfloat16, "gemm", steps=20),
    "cfg2_10Mx768_B24_gemm": lambda: run_search("cfg2_10Mx768_B24_gemm", 10_000_000, 768, 24, 100, torch.bfloat16, "gemm", steps=20),
    "cfg2_10Mx768_B64_gemm": lambda: run_search("cfg2_10Mx768_B64_gemm", 10_000_000, 768, 64, 100, torch.bfloat16, "gemm", steps=20),
    "cfg2_10Mx768_B128_gemm": lambda: run_search("cfg2_10Mx768_B128_gemm", 10_000_000, 768, 128, 100, torch.bfloat16, "gemm", steps=20),
    "cfg2_10Mx768_B16_scan": lambda: run_search("cfg2_10Mx768_B16_scan", 10_000_000, 768, 16, 100, torch.bfloat16, "scan", steps=5),
    "cfg2_10Mx768_B2_scan": lambda: run_search("cfg2_10Mx768_B2_scan", 10_000_000, 768, 2, 100, torch.bfloat16, "scan", steps=20, check=False),
    "cfg2_10Mx768_B2_gemm": lambda: run_search("cfg2_10Mx768_B2_gemm", 10_000_000, 768, 2, 100, torch.bfloat16, "gemm", steps=20, check=False),
    "f32_5Mx768_B1_tensor": lambda: run_search("f32_5Mx768_B1_tensor", 5_000_000, 768, 1, 100, torch.float32, "gemm", steps=20),
    "f32_5Mx768_B8_tensor": lambda: run_search("f32_5Mx768_B8_tensor", 5_000_000, 768, 8, 100, torch.float32, "gemm", steps=20),
    "f32_5Mx768_B16_tensor": lambda: run_search("f32_5Mx768_B16_tensor", 5_000_000, 768, 16, 100, torch.float32, "gemm", steps=20),
    "f32_5Mx768_B64_tensor": lambda: run_search("f32_5Mx768_B64_tensor", 5_000_000, 768, 64, 100, torch.float32, "gemm", steps=10),
    "f32_5Mx768_B256_tensor": lambda: run_search("f32_5Mx768_B256_tensor", 5_000_000, 768, 256, 100, torch.float32, "gemm", steps=5),
    "f32_5Mx768_B1_scan": lambda: run_search("f32_5Mx768_B1_scan", 5_000_000, 768, 1, 100, torch.float32, "scan", steps=20),
    "f32_5Mx768_B4_scan": lambda: run_search("f32_5Mx768_B4_scan", 5_000_000, 768, 4, 100, torch.float32, "scan", steps=20),
    "f32_5Mx768_B8_scan": lambda: run_search("f32_5Mx768_B8_scan", 5_000_000, 768, 8, 100, torch.float32, "scan", steps=20),
    "f32_5Mx768_B16_scan": lambda: run_search("f32_5Mx768_B16_scan", 5_000_000, 768, 16, 100, torch.float32, "scan", steps=10),
    "i8_10Mx768_B1_scan": lambda: run_search_i8("i8_10Mx768_B1_scan", 10_000_000, 768, 1, 100, steps=20, path="scan"),
    "i8_10Mx768_B4_scan": lambda: run_search_i8("i8_10Mx768_B4_scan", 10_000_000, 768, 4, 100, steps=20, path="scan"),
    "i8_10Mx768_B1_tensor": lambda: run_search_i8("i8_10Mx768_B1_tensor", 10_000_000, 768, 1, 100, steps=20),
    "i8_10Mx768_B4_tensor": lambda: run_search_i8("i8_10Mx768_B4_tensor", 10_000_000, 768, 4, 100, steps=20),
    "i8_10Mx768_B32_tensor": lambda: run_search_i8("i8_10Mx768_B32_tensor", 10_000_000, 768, 32, 100, steps=10),
    "i8_10Mx768_B256_tensor": lambda: run_search_i8("i8_10Mx768_B256_tensor", 10_000_000, 768, 256, 100, steps=5),
    "i8_10Mx768_B1024_tensor": lambda: run_search_i8("i8_10Mx768_B1024_tensor", 10_000_000, 768, 1024, 100, steps=3, warmup=2),
    "i8_10Mx768_B4096_tensor": lambda: run_search_i8("i8_10Mx768_B4096_tensor", 10_000_000, 768, 4096, 100, steps=3, warmup=2),
    "i8_10Mx768_B64_tensor": lambda: run_search_i8("i8_10Mx768_B64_tensor", 10_000_000, 768, 64, 100, steps=10),
    "i8_10Mx768_B16_tensor": lambda: run_search_i8("i8_10Mx768_B16_tensor", 10_000_000, 768, 16, 100, steps=20),
    "cfg2_10Mx768_B1_gemm": lambda: run_search("cfg2_10Mx768_B1_gemm", 10_000_000, 768, 1, 100, torch.bfloat16, "gemm", steps=20),
    "cfg3_12p5Mx384_f16_k1000_B16": lambda: run_search("cfg3_12p5Mx384_f16_k1000_B16", 12_500_000, 384, 16, 1000, torch.float16, "gemm", steps=10),
    "cfg3_12p5Mx384_f16_k1000_B1": lambda: run_search("cfg3_12p5Mx384_f16_k1000_B1", 12_500_000, 384, 1, 1000, torch.float16, "scan", steps=10),
    "cfg4_graph_5M_50M": lambda: run_graph("cfg4_graph_5M_50M", 5_000_000, 25_000_000, 768, 1024, 20, 10),
    "cfg4_graph_5M_50M_B4096": lambda: run_graph("cfg4_graph_5M_50M_B4096", 5_000_000, 25_000_000, 768, 4096, 20, 10),
    "cfg4_gather_only_B4096": lambda: run_graph("cfg4_gather_only_B4096", 5_000_000, 2_500_000, 768, 4096, 2200, 10, window=0, random_seeds=True),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        CONFIGS[n]()
