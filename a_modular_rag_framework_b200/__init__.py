"""B200-native retrieval hot path of A-Modular-RAG-Framework: exact cosine/IP top-k over a resident
passage-embedding corpus (tcgen05 GEMM + fused top-k, streaming scan), graph 2-hop expansion +
re-score, and the shard merge -- behind the reference's retriever plug-in interface.

Layout:  csrc/ (CUDA kernels + C ABI, built to _lib/libb200r.so by build.py) . _native (ctypes) .
index (DenseIndex / PassageGraph / expand_rescore / merge_topk) . sharded (row-sharded multi-GPU) .
backend (B200RetrievalBackend, the drop-in for HybridRetrievalBackend) . store (on-disk formats).
"""
__version__ = "0.1.0"

from ._native import B200RError  # noqa: F401


def __getattr__(name):
    # heavy imports (torch) are deferred so that `import a_modular_rag_framework_b200` stays cheap
    if name in ("DenseIndex", "PassageGraph", "expand_rescore", "merge_topk"):
        from . import index as _index
        return getattr(_index, name)
    if name in ("ShardedDenseIndex",):
        from . import sharded as _sharded
        return getattr(_sharded, name)
    if name in ("B200RetrievalBackend",):
        from . import backend as _backend
        return getattr(_backend, name)
    raise AttributeError(name)
