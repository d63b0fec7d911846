"""Row-sharded exact search across the GPUs of one box: one process per GPU (torch.distributed).

Rows of the corpus are independent, and the top-k of a union is the top-k of the per-shard top-k's
(SURVEY.md section 8e), so the path has ONE exchange step: every rank searches its contiguous row
block, the packed (ids, scores)[B, k] results are all-gathered (NCCL over NVLink on the GPU box,
gloo in the CPU tests) and every rank runs the same on-device merge (b200r_merge_topk_packed), ties
resolved by (score desc, GLOBAL id asc) exactly like the single-GPU path.

The local search engine and the merge are injectable so the sharding / exchange logic can be tested
on CPU ranks; the defaults are the CUDA kernels and raise without them (no CPU fallback).
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(N: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) owned by `rank`: ceil(N / world) rows each, the tail may be short/empty."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


def _cuda_merge_packed(packed: torch.Tensor, G: int, B: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    from . import _native as N
    from .index import _stream_ptr
    dev = packed.device
    scores = torch.empty((B, k), dtype=torch.float32, device=dev)
    ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.load().b200r_merge_topk_packed(dev.index or 0, packed.data_ptr(), G, B, k, k, scores.data_ptr(),
                                                 ids.data_ptr(), _stream_ptr(dev)))
    return scores, ids


class ShardedDenseIndex:
    """`local` searches this rank's rows and reports GLOBAL ids (DenseIndex(X_shard, id_base=lo))."""

    def __init__(self, local, group: Optional[dist.ProcessGroup] = None,
                 merge_packed: Optional[Callable[[torch.Tensor, int, int, int], Tuple[torch.Tensor, torch.Tensor]]] = None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._merge = merge_packed or _cuda_merge_packed
        self._pack: Optional[torch.Tensor] = None
        self._gather: Optional[torch.Tensor] = None

    def _buffers(self, B: int, k: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        n = B * k
        if n % 2:
            raise ValueError("B*k must be even for the packed exchange")
        rec = 12 * n
        if self._pack is None or self._pack.numel() != rec or self._pack.device != device:
            self._pack = torch.empty(rec, dtype=torch.uint8, device=device)
            self._gather = torch.empty(rec * self.world, dtype=torch.uint8, device=device)
        return self._pack, self._gather

    def search(self, Q: torch.Tensor, k: int, path: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
        """Every rank passes the SAME Q [B, D]; every rank returns the same global (scores, ids) [B, k]."""
        if self.world == 1:
            return self.local.search(Q, k, path=path)
        B = int(Q.shape[0])
        pack, gather = self._buffers(B, k, Q.device)
        n = B * k
        ids_view = pack[: 8 * n].view(torch.int64).view(B, k)
        sc_view = pack[8 * n:].view(torch.float32).view(B, k)
        self.local.search(Q, k, path=path, out=(sc_view, ids_view))
        dist.all_gather_into_tensor(gather, pack, group=self.group)
        return self._merge(gather, self.world, B, k)
