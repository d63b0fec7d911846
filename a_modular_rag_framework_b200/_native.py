"""ctypes binding of libb200r.so -- the only way the package reaches the GPU.

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

from . import build as _build

BF16, F16, F32, I8 = 0, 1, 2, 3
IP, COS = 0, 1
PATH_AUTO, PATH_GEMM, PATH_SCAN = 0, 1, 2
MAX_K = 2048
EOVERFLOW = -4
ETIMEOUT = -6
COMM_BLOB_BYTES = 128

#: every symbol include/b200r.h declares (tests check the .so exports all of them)
EXPORTS = [
    "b200r_version", "b200r_last_error", "b200r_launch_count",
    "b200r_index_create", "b200r_index_destroy", "b200r_index_info",
    "b200r_search_workspace_bytes", "b200r_search", "b200r_search_host", "b200r_search_host_f32", "b200r_convert_f32", "b200r_search_min_score",
    "b200r_range_workspace_bytes", "b200r_range_count", "b200r_range_fill",
    "b200r_comm_create", "b200r_comm_export", "b200r_comm_connect", "b200r_comm_init_all", "b200r_comm_destroy", "b200r_comm_status", "b200r_comm_info",
    "b200r_search_sharded_begin", "b200r_search_sharded_finish", "b200r_search_sharded", "b200r_search_sharded_host_all", "b200r_search_sharded_host_all_f32",
    "b200r_merge_topk", "b200r_merge_topk_packed", "b200r_mmr_diversify", "b200r_profile_enable", "b200r_profile_read", "b200r_profile_dump",
    "b200r_graph_create", "b200r_graph_create_mapped", "b200r_graph_destroy",
    "b200r_merge_topk_src", "b200r_expand_max_candidates", "b200r_expand_rescore", "b200r_expand_workspace_bytes", "b200r_expand_rescore_sharded", "b200r_expand_rescore_exchange",
    "b200r_expand_big_workspace_bytes", "b200r_expand_rescore_big",
]


class B200RError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200r error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib_path() -> str:
    # B200R_LIB points at an alternative build of the same ABI (A/B measurements of kernel variants)
    return os.environ.get("B200R_LIB") or _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (once) the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path) and not os.environ.get("B200R_LIB"):
        try:
            _build.build()       # in-tree nvcc build; still no CPU fallback if this fails
        except Exception as e:  # noqa: BLE001
            raise B200RError(-2, f"{path} is missing and could not be built ({e}); run "
                                 f"`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback")
    if not os.path.exists(path):
        raise B200RError(-2, f"{path} is missing; there is no CPU fallback")
    lib = ctypes.CDLL(path)
    lib.b200r_version.restype = c_int
    lib.b200r_last_error.argtypes = [c_char_p, c_size_t]
    lib.b200r_last_error.restype = c_int
    lib.b200r_launch_count.restype = c_int64
    lib.b200r_index_create.argtypes = [c_int, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p, POINTER(c_uint64)]
    lib.b200r_index_destroy.argtypes = [c_uint64]
    lib.b200r_index_info.argtypes = [c_uint64, POINTER(c_int64), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int64)]
    lib.b200r_search_workspace_bytes.argtypes = [c_uint64, c_int32, c_int32, c_int32]
    lib.b200r_search_workspace_bytes.restype = c_size_t
    lib.b200r_search.argtypes = [c_uint64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.b200r_search_host.argtypes = [c_uint64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.b200r_search_host_f32.argtypes = [c_uint64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.b200r_convert_f32.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int32, c_void_p]
    lib.b200r_search_min_score.argtypes = [c_uint64, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.b200r_range_workspace_bytes.argtypes = [c_uint64, c_int32]
    lib.b200r_range_workspace_bytes.restype = c_size_t
    lib.b200r_range_count.argtypes = [c_uint64, c_void_p, c_int32, c_float, c_void_p, c_void_p, POINTER(c_int64), c_void_p, c_size_t, c_void_p]
    lib.b200r_range_fill.argtypes = [c_uint64, c_void_p, c_int32, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.b200r_comm_create.argtypes = [c_int, c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(c_uint64)]
    lib.b200r_comm_export.argtypes = [c_uint64, c_void_p]
    lib.b200r_comm_connect.argtypes = [c_uint64, c_void_p]
    lib.b200r_comm_init_all.argtypes = [POINTER(c_int32), c_int32, c_int32, c_int32, c_int32, POINTER(c_uint64)]
    lib.b200r_comm_destroy.argtypes = [c_uint64]
    lib.b200r_comm_status.argtypes = [c_uint64, c_void_p]
    lib.b200r_comm_info.argtypes = [c_uint64, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_size_t)]
    lib.b200r_search_sharded_begin.argtypes = [c_uint64, c_uint64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]
    lib.b200r_search_sharded_finish.argtypes = [c_uint64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.b200r_search_sharded.argtypes = [c_uint64, c_uint64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.b200r_search_sharded_host_all.argtypes = [POINTER(c_uint64), POINTER(c_uint64), c_int32, c_void_p, c_int32, c_int32, c_int32,
                                                  c_void_p, c_void_p, POINTER(c_void_p)]
    lib.b200r_search_sharded_host_all_f32.argtypes = lib.b200r_search_sharded_host_all.argtypes
    lib.b200r_mmr_diversify.argtypes = [c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.b200r_merge_topk.argtypes = [c_int, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.b200r_merge_topk_packed.argtypes = [c_int, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.b200r_merge_topk_src.argtypes = [c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_int32, c_int32, c_int32, c_int32,
                                         c_void_p, c_void_p, c_void_p, c_void_p]
    lib.b200r_profile_enable.argtypes = [c_int]
    lib.b200r_profile_read.argtypes = [POINTER(ctypes.c_double), POINTER(c_int64)]
    lib.b200r_graph_create.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int64, POINTER(c_uint64)]
    lib.b200r_graph_create_mapped.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, POINTER(c_uint64)]
    lib.b200r_graph_destroy.argtypes = [c_uint64]
    lib.b200r_expand_max_candidates.restype = c_int32
    lib.b200r_expand_rescore.argtypes = [c_uint64, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_int32,
                                         c_void_p, c_size_t,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.b200r_expand_workspace_bytes.argtypes = [c_int32]
    lib.b200r_expand_workspace_bytes.restype = c_size_t
    lib.b200r_expand_rescore_sharded.argtypes = [c_uint64, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_int32,
                                                 c_int32, c_int64, c_int64, c_void_p, c_size_t, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.b200r_expand_big_workspace_bytes.argtypes = [c_uint64]
    lib.b200r_expand_big_workspace_bytes.restype = c_size_t
    lib.b200r_expand_rescore_big.argtypes = [c_uint64, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_int32, c_void_p, c_size_t,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.b200r_expand_rescore_exchange.argtypes = [c_uint64, c_uint64, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float,
                                                  c_int32, c_int64, c_int32, c_void_p, c_size_t,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is c_int and name not in ("b200r_version", "b200r_last_error"):
            fn.restype = c_int
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().b200r_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise B200RError(rc, last_error())


def launch_count() -> int:
    return int(load().b200r_launch_count())


def profile_enable(on: bool) -> None:
    check(load().b200r_profile_enable(1 if on else 0))


def profile_timeline(on: bool) -> None:
    """Bracket every launch site with events (profile_dump reads them)."""
    check(load().b200r_profile_enable(2 if on else 0))


def profile_dump() -> dict:
    import json
    lib = load()
    lib.b200r_profile_dump.argtypes = [c_char_p, c_size_t]
    buf = ctypes.create_string_buffer(1 << 16)
    lib.b200r_profile_dump(buf, 1 << 16)
    return json.loads(buf.value.decode() or "{}")


def profile_read():
    """-> (summed device ms of the dominant search kernel since the last read, number of launches)"""
    ms, n = ctypes.c_double(0.0), c_int64(0)
    check(load().b200r_profile_read(ctypes.byref(ms), ctypes.byref(n)))
    return float(ms.value), int(n.value)
