"""dev helper (not part of the product): quick GPU sanity of each path with verbose diagnostics."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import retrieval_oracle as O
from a_modular_rag_framework_b200 import DenseIndex

dev = torch.device("cuda:0")
def run(N, D, B, k, path, dtype=torch.bfloat16, metric="cos"):
    g = torch.Generator(device="cuda").manual_seed(N + D + B)
    X = torch.randn((N, D), generator=g, device=dev).to(dtype)
    Q = torch.randn((B, D), generator=g, device=dev).to(dtype)
    ix = DenseIndex(X, metric=metric)
    t0 = time.time()
    s, i = ix.search(Q, k, path=path)
    torch.cuda.synchronize()
    dt = time.time() - t0
    S = O.score_matrix(Q.float().cpu().numpy(), X.float().cpu().numpy(), metric)
    rs, ri = O.topk_from_scores(S, k)
    i = i.cpu().numpy(); s = s.cpu().numpy()
    kk = min(k, N)
    match = (i[:, :kk] == ri[:, :kk]).mean()
    err = np.abs(s[:, :kk] - rs[:, :kk]).max()
    print(f"N={N} D={D} B={B} k={k} path={path} {dtype}: id-match={match:.4f} max|ds|={err:.2e} t={dt*1e3:.1f}ms", flush=True)
    if match < 0.99:
        print(" got ids", i[0, :10], "\n ref ids", ri[0, :10], "\n got s", s[0, :10], "\n ref s", rs[0, :10])

if __name__ == "__main__":
    run(2000, 64, 2, 10, "scan")
    run(20000, 128, 4, 100, "scan")
    run(2000, 64, 8, 10, "gemm")
    run(20000, 128, 37, 100, "gemm")
    run(100000, 768, 256, 100, "gemm")
