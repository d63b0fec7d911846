"""ShardComm -- host handle of the C-ABI shard exchange (include/b200r.h, b200r_comm_*).

The data plane is inside libb200r.so: kernels store into the peers' windows over NVLink and wait on device-side
epoch flags (csrc/comm.cuh).  What remains for the host is the BRING-UP: every rank exports a 128-byte descriptor of
its window and needs everybody else's.  Any byte channel does; three are provided:

  * ``ShardComm.init_all(devices, ...)``            one process driving several GPUs (no channel needed)
  * ``ShardComm.from_files(dir, rank, world, ...)`` a shared directory (ctypes-only hosts, no torch.distributed)
  * ``ShardComm.from_process_group(group, ...)``    torch.distributed's object all-gather (control plane only)

The reference has no counterpart (it is one process with one in-memory index, retrieval_backend.py:252-281); this is
the SURVEY.md section 8(b) `b200r_comm_init_all` / `b200r_allgather_topk` requirement, done with peer stores
instead of an NCCL collective.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import List, Optional, Sequence, Tuple

from . import _native as N


class ShardComm:
    def __init__(self, handle: int, rank: int, world: int, device: int, owner: bool = True):
        self._lib = N.load()
        self._handle = ctypes.c_uint64(handle)
        self.rank, self.world, self.device_index = int(rank), int(world), int(device)
        self._owner = owner

    # -- bring-up ---------------------------------------------------------------------------------------
    @classmethod
    def create(cls, device: int, rank: int, world: int, max_B: int, max_k: int, max_graph_k: int = 0) -> "ShardComm":
        lib = N.load()
        h = ctypes.c_uint64(0)
        N.check(lib.b200r_comm_create(int(device), int(rank), int(world), int(max_B), int(max_k), int(max_graph_k), ctypes.byref(h)))
        return cls(h.value, rank, world, device)

    def export(self) -> bytes:
        buf = ctypes.create_string_buffer(N.COMM_BLOB_BYTES)
        N.check(self._lib.b200r_comm_export(self._handle, buf))
        return buf.raw

    def connect(self, blobs: Sequence[bytes]) -> None:
        if len(blobs) != self.world or any(len(b) != N.COMM_BLOB_BYTES for b in blobs):
            raise ValueError(f"need {self.world} descriptors of {N.COMM_BLOB_BYTES} bytes, in rank order")
        N.check(self._lib.b200r_comm_connect(self._handle, b"".join(blobs)))

    @classmethod
    def init_all(cls, devices: Sequence[int], max_B: int, max_k: int, max_graph_k: int = 0) -> List["ShardComm"]:
        """One process, G GPUs: comms[g] lives on devices[g] and is rank g of G."""
        lib = N.load()
        G = len(devices)
        devs = (ctypes.c_int32 * G)(*[int(d) for d in devices])
        out = (ctypes.c_uint64 * G)()
        N.check(lib.b200r_comm_init_all(devs, G, int(max_B), int(max_k), int(max_graph_k), out))
        return [cls(out[g], g, G, devices[g]) for g in range(G)]

    @classmethod
    def from_files(cls, directory: str, device: int, rank: int, world: int, max_B: int, max_k: int, max_graph_k: int = 0,
                   timeout_s: float = 120.0, tag: str = "comm") -> "ShardComm":
        """Descriptor exchange through a directory every rank can see (no torch.distributed, no sockets)."""
        c = cls.create(device, rank, world, max_B, max_k, max_graph_k)
        if world == 1:
            return c
        os.makedirs(directory, exist_ok=True)
        mine = os.path.join(directory, f"{tag}.{rank}.blob")
        with open(mine + ".tmp", "wb") as f:
            f.write(c.export())
        os.replace(mine + ".tmp", mine)
        blobs, t0 = [], time.time()
        for g in range(world):
            path = os.path.join(directory, f"{tag}.{g}.blob")
            while not os.path.exists(path):
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"rank {g} never published {path}")
                time.sleep(0.01)
            with open(path, "rb") as f:
                blobs.append(f.read())
        c.connect(blobs)
        return c

    @classmethod
    def from_process_group(cls, device: int, max_B: int, max_k: int, max_graph_k: int = 0, group=None) -> "ShardComm":
        """Descriptor exchange through torch.distributed (any backend; control plane only)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = cls.create(device, rank, world, max_B, max_k, max_graph_k)
        if world > 1:
            blobs: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(blobs, c.export(), group=group)
            err = None
            try:
                c.connect(blobs)         # type: ignore[arg-type]
            except N.B200RError as e:
                err = f"rank {rank}: {e}"
            # agree on the outcome (this also orders "every rank has mapped every window" before the first store)
            errs: List[Optional[str]] = [None] * world
            dist.all_gather_object(errs, err, group=group)
            bad = [e for e in errs if e]
            if bad:
                c.close()
                raise N.B200RError(-2, "exchange bring-up failed: " + "; ".join(bad))
        return c

    # -- data plane ---------------------------------------------------------------------------------------
    @property
    def handle(self) -> int:
        if not self._handle.value:
            raise RuntimeError("comm is closed")
        return self._handle.value

    def finish(self, B: int, k: int, out: Optional[Tuple["torch.Tensor", "torch.Tensor"]] = None):   # noqa: F821
        """Receiving half of DenseIndex.search_sharded_begin: one kernel that waits for every rank's record and merges
        the sorted per-shard lists -> (scores f32 [B, k], ids i64 [B, k]) on this rank's device.  Stream-ordered."""
        import torch
        from .index import _stream_ptr
        dev = torch.device("cuda", self.device_index)
        if out is None:
            scores = torch.empty((B, k), dtype=torch.float32, device=dev)
            ids = torch.empty((B, k), dtype=torch.int64, device=dev)
        else:
            scores, ids = out
        with torch.cuda.device(dev):
            N.check(self._lib.b200r_search_sharded_finish(self._handle, int(B), int(k), scores.data_ptr(), ids.data_ptr(), _stream_ptr(dev)))
        return scores, ids

    def check(self, stream_ptr: int = 0) -> None:
        """Raises B200RError(ETIMEOUT) if a device-side wait of an earlier exchange expired (synchronises the stream)."""
        N.check(self._lib.b200r_comm_status(self._handle, stream_ptr or None))

    def window_bytes(self) -> int:
        wb = ctypes.c_size_t(0)
        N.check(self._lib.b200r_comm_info(self._handle, None, None, None, ctypes.byref(wb)))
        return int(wb.value)

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value and self._owner:
            self._lib.b200r_comm_destroy(self._handle)
        self._handle = ctypes.c_uint64(0)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
