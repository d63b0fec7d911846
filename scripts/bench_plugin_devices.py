"""The plug-in engine on ALL visible GPUs from ONE process (what B200RetrievalBackend(devices=[...]) builds): float32 host
queries in, host results out, corpus row-sharded over the devices (b200r_search_sharded_host_all_f32).  Reports the
host-to-host latency of one call at batch 1 / 16 / 4096 -- batch 1 is how the reference's orchestrator calls its retriever
(app/orchestrator/nodes.py:110-115).

    python scripts/bench_plugin_devices.py [--rows 10000000] [--dim 768] [--k 100]
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_modular_rag_framework_b200.backend import CudaEngine
from a_modular_rag_framework_b200.sharded import shard_rows
from a_modular_rag_framework_b200 import DenseIndex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=100)
    a = ap.parse_args()
    G = torch.cuda.device_count()
    shards = []
    for g in range(G):
        lo, hi = shard_rows(a.rows, G, g)
        dev = torch.device("cuda", g)
        X = torch.empty((hi - lo, a.dim), dtype=torch.bfloat16, device=dev)
        for r0 in range(0, hi - lo, 1 << 20):
            gen = torch.Generator(device=dev).manual_seed(1000 + lo + r0)
            t = torch.randn((min(1 << 20, hi - lo - r0), a.dim), generator=gen, device=dev)
            X[r0:r0 + t.shape[0]] = (t / t.norm(dim=1, keepdim=True)).to(torch.bfloat16)
        shards.append(X)
    eng = CudaEngine.from_shards(shards, a.rows, "cos", max_batch=4096, max_k=a.k)
    rng = np.random.default_rng(0)
    out = {"devices": G, "rows": a.rows, "dim": a.dim, "k": a.k}
    for B, reps in ((1, 300), (16, 300), (4096, 10)):
        Qs = [rng.standard_normal((B, a.dim)).astype(np.float32) for _ in range(4)]
        for i in range(3):
            eng.dense_search(Qs[i % 4], a.k)
        t0 = time.perf_counter()
        for i in range(reps):
            s, ids = eng.dense_search(Qs[i % 4], a.k)
        ms = (time.perf_counter() - t0) / reps * 1e3
        out[f"batch{B}_ms_per_call"] = ms
        out[f"batch{B}_queries_per_s"] = B / (ms * 1e-3)
    # the last batch-4096 result against ONE index over the first shard's device, for the rows that device owns (sanity)
    q = torch.from_numpy(Qs[(reps - 1) % 4][:4]).to(shards[0].device).to(torch.bfloat16)
    best = None
    for g in range(G):
        lo, _ = shard_rows(a.rows, G, g)
        sg, ig = DenseIndex(shards[g], metric="cos", id_base=lo).search(q.to(shards[g].device), a.k)
        cat_s = sg.cpu() if best is None else torch.cat([best[0], sg.cpu()], 1)
        cat_i = ig.cpu() if best is None else torch.cat([best[1], ig.cpu()], 1)
        best = (cat_s, cat_i)
    order = torch.argsort(best[0], dim=1, descending=True, stable=True)[:, :a.k]
    out["ids_equal_per_shard_merge_on_4_queries"] = bool((torch.gather(best[1], 1, order) == torch.from_numpy(ids[:4])).all().item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
