"""Build recipe for libb200r.so (hand-written sm_100a CUDA behind the C ABI of include/b200r.h).

In-tree, explicit nvcc: the .so travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200r.so")
SOURCES = ["api.cu", "gemm_topk.cu", "scan_topk.cu", "scan_bulk.cu", "rows_i8.cu", "rows_f32.cu", "select_merge.cu", "csr_expand.cu", "comm.cu", "mmr.cu"]
HEADERS = ["common.cuh", "comm.cuh", "internal.h", "tc_ptx.cuh", os.path.join("..", "..", "include", "b200r.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(LIB_DIR, "obj")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def _deps_mtime() -> float:
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps if os.path.exists(d))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return _deps_mtime() > t or any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a (one nvcc per translation unit, in parallel; objects whose source and headers are
    unchanged are reused) and link _lib/libb200r.so; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hdr_t = _nvcc(), _deps_mtime()
    jobs, objs = [], []
    for src in SOURCES:
        spath, obj = os.path.join(CSRC, src), os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(spath)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, spath]
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} ---\n{out}")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libb200r.so")
    proc = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("link of libb200r.so failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
