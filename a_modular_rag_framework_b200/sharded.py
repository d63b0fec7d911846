"""Row-sharded exact search across the GPUs of one box: one process per GPU.

Rows of the corpus are independent, and the top-k of a union is the top-k of the per-shard top-k's
(SURVEY.md section 8e), so the path has ONE exchange step: every rank searches its contiguous row
block, every rank's packed (ids, scores)[B, k] record reaches every rank, and every rank runs the same
on-device merge, ties resolved by (score desc, GLOBAL id asc) exactly like the single-GPU path.

On the GPU box the exchange is the library's own (exchange="abi": b200r_comm_*, peer stores over NVLink from
inside the final select kernel, device-side epoch flags; torch.distributed only swaps the 128-byte window
descriptors at bring-up).  exchange="nccl" keeps the packed all-gather + b200r_merge_topk_packed; it is also
what the gloo CPU tests drive, with an injected local engine and merge.  There is no CPU fallback: the
defaults are the CUDA kernels and raise without them.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Callable, Optional, Tuple

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


class PendingResult:
    """Result of ShardedDenseIndex.search(..., overlap=True).  get() returns (scores, ids) on the caller's current stream;
    the receiving half of the exchange is issued either by the NEXT search call (after its local search has been queued,
    so no rank ever waits for a slower peer) or by get(), whichever comes first."""

    def __init__(self, scores=None, ids=None, event=None, finish=None):
        self._scores, self._ids, self._event, self._finish = scores, ids, event, finish

    def resolve(self):
        if self._finish is not None:
            fin, self._finish = self._finish, None
            self._scores, self._ids = fin()
            # get() may be called under another stream (e.g. a copy stream): order it behind the merge kernel
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(self._scores.device))

    def get(self):
        self.resolve()
        if self._event is not None:
            torch.cuda.current_stream(self._scores.device).wait_event(self._event)
            self._event = None
        return self._scores, self._ids


class ShardedDenseIndex:
    """`local` searches this rank's rows and reports GLOBAL ids (DenseIndex(X_shard, id_base=lo)).

    exchange="abi" (what "auto" picks for a CUDA DenseIndex): the C-ABI exchange (comm.ShardComm): the final select kernel of the local search stores
    this rank's packed record into every rank's window over NVLink and one merge kernel waits for the records -- no
    NCCL, no collective kernel, no staging copy, no side stream.  torch.distributed is used once, to swap the 128-byte
    window descriptors.  exchange="nccl": packed all_gather_into_tensor + b200r_merge_topk_packed (also what the gloo CPU
    tests drive, with an injected engine and merge)."""

    def __init__(self, local, group: Optional[dist.ProcessGroup] = None,
                 merge_packed: Optional[Callable[[torch.Tensor, int, int, int], Tuple[torch.Tensor, torch.Tensor]]] = None,
                 exchange: str = "auto", comm=None):
        if exchange == "p2p":        # round-1 name of the fused exchange
            exchange = "abi"
        if exchange == "auto":       # the library's exchange whenever the local engine is a CUDA index; collectives otherwise
            dev = getattr(local, "device", None)
            exchange = "abi" if (hasattr(local, "search_sharded_begin") and getattr(dev, "type", None) == "cuda") else "nccl"
        if exchange not in ("nccl", "abi"):
            raise ValueError("exchange must be 'auto', 'abi' or 'nccl'")
        self.exchange = exchange
        self.comm = comm
        self._comm_cap = (0, 0)
        self._outstanding: Optional[PendingResult] = None
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else (comm.world if comm is not None else 1)
        self.rank = dist.get_rank(group) if dist.is_initialized() else (comm.rank if comm is not None else 0)
        self._merge = merge_packed or _cuda_merge_packed
        self._bufs = {}          # (B, k, slot) -> (pack, gather)
        self._slot = 0
        self._comm_stream = None

    def _buffers(self, B: int, k: int, device, slot: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        n = B * k
        if n % 2:
            raise ValueError("B*k must be even for the packed exchange")
        rec = 12 * n
        key = (B, k, slot, str(device))
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(rec, dtype=torch.uint8, device=device),
                               torch.empty(rec * self.world, dtype=torch.uint8, device=device))
        return self._bufs[key]

    def ensure_comm(self, B: int, k: int):
        """(Re)create the exchange window for batches up to (B, k).  Collective: every rank calls it with the same sizes
        (search() does, since every rank issues the same searches)."""
        from .comm import ShardComm
        if self.comm is not None and self._comm_cap == (0, 0):
            return self.comm                      # supplied by the caller: its capacity is the caller's business
        mb, mk = self._comm_cap
        if self.comm is None or B > mb or k > mk:
            if self._outstanding is not None:
                self._outstanding.resolve()
                self._outstanding = None
            if self.comm is not None:
                torch.cuda.synchronize()
                if dist.is_initialized():
                    dist.barrier(group=self.group)   # nobody is still storing into a window that is about to go away
                self.comm.close()
            mb, mk = max(mb, B), max(mk, k)
            dev = self.local.device
            self.comm = ShardComm.from_process_group(dev.index or 0, mb, mk, 0, group=self.group)
            self._comm_cap = (mb, mk)
        return self.comm

    def search(self, Q: torch.Tensor, k: int, path: str = "auto", overlap: bool = False):
        """Every rank passes the SAME Q [B, D]; every rank gets the same global (scores, ids) [B, k].

        overlap=True (throughput mode): returns a PendingResult; the receiving half of this batch's exchange is issued
        after the NEXT batch's local search has been queued."""
        if self.world == 1:
            res = self.local.search(Q, k, path=path)
            return PendingResult(res[0], res[1], None) if overlap else res
        B = int(Q.shape[0])
        n = B * k
        if self.exchange == "abi":
            return self._search_abi(Q, k, path, overlap)
        if not overlap:
            pack, gather = self._buffers(B, k, Q.device)
            ids_view = pack[: 8 * n].view(torch.int64).view(B, k)
            sc_view = pack[8 * n:].view(torch.float32).view(B, k)
            self.local.search(Q, k, path=path, out=(sc_view, ids_view))
            dist.all_gather_into_tensor(gather, pack, group=self.group)
            return self._merge(gather, self.world, B, k)
        # double-buffered exchange on a side stream
        self._slot ^= 1
        pack, gather = self._buffers(B, k, Q.device, self._slot)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=Q.device)
        cur = torch.cuda.current_stream(Q.device)
        # the buffers of this slot were last used two batches ago on the side stream
        ids_view = pack[: 8 * n].view(torch.int64).view(B, k)
        sc_view = pack[8 * n:].view(torch.float32).view(B, k)
        prev = getattr(self, "_slot_events", {}).get(self._slot)
        if prev is not None:
            cur.wait_event(prev)          # do not overwrite `pack` before its previous all-gather has read it
        self.local.search(Q, k, path=path, out=(sc_view, ids_view))
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ready)
            dist.all_gather_into_tensor(gather, pack, group=self.group)
            scores, ids = self._merge(gather, self.world, B, k)
            done = torch.cuda.Event()
            done.record(self._comm_stream)
        if not hasattr(self, "_slot_events"):
            self._slot_events = {}
        self._slot_events[self._slot] = done
        return PendingResult(scores, ids, done)

    # ---- C-ABI exchange over peer memory ------------------------------------------------------------------
    def _search_abi(self, Q: torch.Tensor, k: int, path: str, overlap: bool):
        B = int(Q.shape[0])
        comm = self.ensure_comm(B, k)
        self.local.search_sharded_begin(comm, Q, k, path=path)
        prev, self._outstanding = self._outstanding, None
        if prev is not None:
            prev.resolve()                # finish(i-1) is queued behind begin(i): its peers arrived a batch ago
        if not overlap:
            return comm.finish(B, k)
        cur = PendingResult(finish=lambda: comm.finish(B, k))
        self._outstanding = cur
        return cur


INT64_MIN, INT64_MAX = -(1 << 63), (1 << 63) - 1


def owned_range(N: int, world: int, rank: int) -> Tuple[int, int]:
    """Global node ids whose candidates `rank` scores in the two-pass owner-computes graph stage
    (b200r_expand_rescore_sharded): its row block; the first rank also takes ids < 0 (none) and the last rank everything
    >= its block start (graph nodes that are not corpus rows score 0.0, exactly as on one GPU)."""
    lo, hi = shard_rows(N, world, rank)
    return (INT64_MIN if rank == 0 else lo), (INT64_MAX if rank == world - 1 else hi)


class ShardedGraphStage:
    """Graph 2-hop expansion + re-score over a ROW-SHARDED corpus, one rank per process, entirely over
    the library's exchange (b200r_expand_rescore_exchange): the CSR is replicated, queries are sharded by HOME rank (BFS,
    fusion, top-k), candidates by the rank that OWNS their rows (gather + re-score), and every transfer is a peer store from
    inside the kernels -- no NCCL collective and no host-side reduction; torch.distributed only swaps the window descriptors."""

    def __init__(self, local_index, graph, N_total: int, group: Optional[dist.ProcessGroup] = None, comm=None,
                 max_B: int = 4096, max_k: int = 256):
        self.index, self.graph, self.N_total, self.group = local_index, graph, int(N_total), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.comm, self._cap = comm, (int(max_B), int(max_k))

    def run(self, Q: torch.Tensor, seeds: torch.Tensor, window: int, alpha_dense: float, alpha_graph: float, k: int):
        from .comm import ShardComm
        from .index import expand_rescore, expand_rescore_exchange
        if self.world == 1:
            return expand_rescore(self.index, self.graph, Q, seeds, window, alpha_dense, alpha_graph, k, check_status=False)
        B = int(Q.shape[0])
        if self.comm is None:
            mb, mk = max(self._cap[0], B), max(self._cap[1], k)
            self.comm = ShardComm.from_process_group(self.index.device.index or 0, mb, mk, max_graph_k=mk, group=self.group)
        return expand_rescore_exchange(self.index, self.graph, self.comm, Q, seeds, window, alpha_dense, alpha_graph, k, self.N_total)
