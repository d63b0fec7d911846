"""Host-side mirrors of the C ABI objects: DenseIndex (corpus), PassageGraph (CSR) and the calls on them.

PyTorch is used for device memory, streams and dtype plumbing only; every computation is a kernel of
libb200r.so reached through ctypes (``_native``).  Nothing here falls back to the CPU.

Reference behaviour replaced (paths relative to the reference root):
  DenseIndex.search        DenseReranker.score + sort/slice   app/modules/retrieval/retrieval_backend.py:215-247, 371-372
  PassageGraph             load_graph_json + build_index      app/modules/retrieval/graph_utils.py:13-55
  expand_rescore           expand_qmatch_neighbors + fusion   graph_utils.py:58-129, retrieval_backend.py:296-301, 353-372
  merge_topk               the same stable-sort rule over per-shard lists
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _native as N

_DTYPES = {torch.bfloat16: N.BF16, torch.float16: N.F16, torch.float32: N.F32, torch.int8: N.I8}
_METRICS = {"ip": N.IP, "cos": N.COS, "cosine": N.COS}
_PATHS = {"auto": N.PATH_AUTO, "gemm": N.PATH_GEMM, "scan": N.PATH_SCAN}


def _require_cuda(t: torch.Tensor, name: str, contiguous: bool = True) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class DenseIndex:
    """Exact cosine / inner-product top-k over a resident [N, D] passage-embedding matrix."""

    def __init__(self, X: torch.Tensor, metric: str = "cos", id_base: int = 0, row_scale: Optional[torch.Tensor] = None):
        _require_cuda(X, "X")
        if X.dim() != 2:
            raise ValueError("X must be [N, D]")
        if X.dtype not in _DTYPES:
            raise ValueError(f"unsupported corpus dtype {X.dtype}")
        if metric not in _METRICS:
            raise ValueError(f"unknown metric {metric!r}")
        if row_scale is not None:
            _require_cuda(row_scale, "row_scale")
            if row_scale.dtype != torch.float32 or row_scale.numel() != X.shape[0]:
                raise ValueError("row_scale must be float32 [N]")
        self._lib = N.load()
        self.X = X                      # borrowed by the library for the index lifetime
        self.row_scale = row_scale
        self.metric = metric
        self.id_base = int(id_base)
        self.device = X.device
        self.N, self.D = int(X.shape[0]), int(X.shape[1])
        self.query_dtype = torch.float32 if X.dtype == torch.int8 else X.dtype
        self._ws: Optional[torch.Tensor] = None
        self._handle = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            N.check(self._lib.b200r_index_create(
                self.device.index or 0, X.data_ptr(), self.N, self.D, _DTYPES[X.dtype], _METRICS[metric], self.id_base,
                row_scale.data_ptr() if row_scale is not None else None, _stream_ptr(self.device), ctypes.byref(self._handle)))

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.b200r_index_destroy(self._handle)
            self._handle = ctypes.c_uint64(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        if not self._handle.value:
            raise RuntimeError("index is closed")
        return self._handle.value

    # -- search -----------------------------------------------------------------------------------
    def _workspace(self, B: int, k: int, path: int) -> torch.Tensor:
        need = int(self._lib.b200r_search_workspace_bytes(self._handle, B, k, path))
        if need == 0:
            raise N.B200RError(-1, N.last_error())
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def search(self, Q: torch.Tensor, k: int, path: str = "auto",
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, min_score: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Q [B, D] on the device in ``query_dtype`` -> (scores f32 [B, k], ids i64 [B, k]), stream-ordered, no sync.
        min_score: only results with score >= min_score are reported (the rest of each row is (-inf, -1))."""
        _require_cuda(Q, "Q")
        if Q.dim() != 2 or Q.shape[1] != self.D:
            raise ValueError(f"Q must be [B, {self.D}]")
        if Q.dtype != self.query_dtype:
            raise ValueError(f"Q must be {self.query_dtype} for a {self.X.dtype} corpus")
        B = int(Q.shape[0])
        p = _PATHS[path]
        if out is None:
            scores = torch.empty((B, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((B, k), dtype=torch.int64, device=self.device)
        else:
            scores, ids = out
        ws = self._workspace(B, k, p)
        with torch.cuda.device(self.device):
            if min_score is None:
                N.check(self._lib.b200r_search(self._handle, Q.data_ptr(), B, k, p, scores.data_ptr(), ids.data_ptr(),
                                               ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
            else:
                N.check(self._lib.b200r_search_min_score(self._handle, Q.data_ptr(), B, k, p, float(min_score), scores.data_ptr(),
                                                         ids.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        return scores, ids

    def search_sharded_begin(self, comm, Q: torch.Tensor, k: int, path: str = "auto") -> None:
        """Local search of this rank's rows whose final select kernel stores the packed (ids | scores)[B, k] record into
        every rank's exchange window (comm.ShardComm).  Stream-ordered, no sync; pair with comm.finish()."""
        _require_cuda(Q, "Q")
        if Q.dim() != 2 or Q.shape[1] != self.D or Q.dtype != self.query_dtype:
            raise ValueError(f"Q must be [B, {self.D}] {self.query_dtype}")
        B = int(Q.shape[0])
        p = _PATHS[path]
        ws = self._workspace(B, k, p)
        with torch.cuda.device(self.device):
            N.check(self._lib.b200r_search_sharded_begin(self._handle, comm.handle, Q.data_ptr(), B, k, p,
                                                         ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))

    def to_query_dtype(self, Q32: torch.Tensor) -> torch.Tensor:
        """fp32 queries on this index's device -> its query dtype, converted by the library (b200r_convert_f32, RNE)."""
        _require_cuda(Q32, "Q32")
        if Q32.dtype != torch.float32 or not Q32.is_contiguous():
            raise ValueError("Q32 must be a contiguous float32 CUDA tensor")
        if self.query_dtype == torch.float32:
            return Q32
        out = torch.empty(Q32.shape, dtype=self.query_dtype, device=Q32.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.b200r_convert_f32(self.device.index or 0, Q32.data_ptr(), out.data_ptr(), Q32.numel(),
                                                _DTYPES[self.query_dtype], _stream_ptr(self.device)))
        return out

    def search_host(self, Q_host: torch.Tensor, k: int, path: str = "auto",
                    out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host buffers in, host buffers out (H2D + search + D2H + stream sync inside the C call).  Q_host is in the index's
        query dtype, or float32 (what an embedding provider returns): the library then rounds it to the query dtype on the device."""
        if Q_host.is_cuda or not Q_host.is_contiguous() or Q_host.dtype not in (self.query_dtype, torch.float32) or Q_host.shape[1] != self.D:
            raise ValueError("Q_host must be a contiguous CPU tensor [B, D] in the index's query dtype or float32")
        call = self._lib.b200r_search_host if Q_host.dtype == self.query_dtype else self._lib.b200r_search_host_f32
        B = int(Q_host.shape[0])
        if out is None:
            scores = torch.empty((B, k), dtype=torch.float32).pin_memory()
            ids = torch.empty((B, k), dtype=torch.int64).pin_memory()
        else:
            scores, ids = out
        with torch.cuda.device(self.device):
            N.check(call(self._handle, Q_host.data_ptr(), B, k, _PATHS[path], scores.data_ptr(), ids.data_ptr(), _stream_ptr(self.device)))
        return scores, ids


def mmr_diversify(index: DenseIndex, cand_ids: torch.Tensor, cand_scores: torch.Tensor, top_k: int, lambda_weight: float = 0.7
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """mmr_diversify (app/utils/similarity.py:33-62) for B candidate lists [B, P] at once; the item vectors are the corpus rows
    behind the ids.  -> (ids i64 [B, top_k], scores f32 [B, top_k], pos i32 [B, top_k]) in pick order."""
    _require_cuda(cand_ids, "cand_ids")
    _require_cuda(cand_scores, "cand_scores")
    if cand_ids.dtype != torch.int64 or cand_scores.dtype != torch.float32 or cand_ids.shape != cand_scores.shape or cand_ids.dim() != 2:
        raise ValueError("cand_ids i64 / cand_scores f32 must both be [B, P]")
    B, P = (int(x) for x in cand_ids.shape)
    dev = index.device
    ids = torch.empty((B, top_k), dtype=torch.int64, device=dev)
    sc = torch.empty((B, top_k), dtype=torch.float32, device=dev)
    pos = torch.empty((B, top_k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.load().b200r_mmr_diversify(index.handle, cand_ids.data_ptr(), cand_scores.data_ptr(), B, P, int(top_k), float(lambda_weight),
                                             ids.data_ptr(), sc.data_ptr(), pos.data_ptr(), _stream_ptr(dev)))
    return ids, sc, pos


class CapturedSearch:
    """One search step (local or row-sharded: search + exchange + merge) captured into a CUDA graph and replayed: a
    batch-1 step is a dozen short kernels, so its latency is the host's launch rate unless they go down as ONE graph
    launch.  libb200r's calls are capture-safe (no allocation, no synchronisation; the exchange keeps its epochs on the
    device).  `index` is a DenseIndex or a ShardedDenseIndex; the query buffer and the outputs are static tensors:
    copy a new batch into `.Q` (or pass it to __call__) and replay.  Every rank of a sharded index must replay in step."""

    def __init__(self, index, Q: torch.Tensor, k: int, path: str = "auto"):
        self.index, self.k, self.path = index, int(k), path
        self.Q = Q.clone()
        index.search(self.Q, self.k, path=path)                   # warm-up outside the capture: workspaces, windows, attributes
        torch.cuda.synchronize(self.Q.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.scores, self.ids = index.search(self.Q, self.k, path=path)

    def __call__(self, Q: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if Q is not None:
            self.Q.copy_(Q, non_blocking=True)
        self.graph.replay()
        return self.scores, self.ids


def merge_topk(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G, B, kin] per-shard lists -> [B, k] by (score desc, id asc); ids < 0 are padding."""
    _require_cuda(cand_scores, "cand_scores")
    _require_cuda(cand_ids, "cand_ids")
    if cand_scores.dtype != torch.float32 or cand_ids.dtype != torch.int64 or cand_scores.shape != cand_ids.shape or cand_scores.dim() != 3:
        raise ValueError("cand_scores f32 / cand_ids i64 must both be [G, B, kin]")
    G, B, kin = (int(x) for x in cand_scores.shape)
    dev = cand_scores.device
    scores = torch.empty((B, k), dtype=torch.float32, device=dev)
    ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.load().b200r_merge_topk(dev.index or 0, cand_scores.data_ptr(), cand_ids.data_ptr(), G, B, kin, k,
                                          scores.data_ptr(), ids.data_ptr(), _stream_ptr(dev)))
    return scores, ids


def merge_topk_src(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Like merge_topk, for [G, B, kin] views whose shard stride (dim 0) may exceed B*kin (slices of one gathered
    buffer), and also returns src [B, k] int32 = g*kin + j, where each winner sits in its shard's list (-1 = padding)."""
    _require_cuda(cand_scores, "cand_scores", contiguous=False)
    _require_cuda(cand_ids, "cand_ids", contiguous=False)
    if cand_scores.dtype != torch.float32 or cand_ids.dtype != torch.int64 or cand_scores.shape != cand_ids.shape or cand_scores.dim() != 3:
        raise ValueError("cand_scores f32 / cand_ids i64 must both be [G, B, kin]")
    G, B, kin = (int(x) for x in cand_scores.shape)
    for t in (cand_scores, cand_ids):
        if t.stride(2) != 1 or t.stride(1) != kin or (G > 1 and t.stride(0) < B * kin):
            raise ValueError("each shard's [B, kin] block must be dense")
    dev = cand_scores.device
    scores = torch.empty((B, k), dtype=torch.float32, device=dev)
    ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    src = torch.empty((B, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.load().b200r_merge_topk_src(dev.index or 0, cand_scores.data_ptr(), max(cand_scores.stride(0), B * kin),
                                              cand_ids.data_ptr(), max(cand_ids.stride(0), B * kin), G, B, kin, k,
                                              scores.data_ptr(), ids.data_ptr(), src.data_ptr(), _stream_ptr(dev)))
    return scores, ids, src


class PassageGraph:
    """CSR of the undirected next_in_doc adjacency.  Without `node_rows` node id == corpus row id; with it the node ids are
    local and node_rows[v] (int64, < 0 = none) is node v's corpus row (per-graph_id graphs, block-diagonal batches)."""

    def __init__(self, row_ptr: torch.Tensor, col_idx: torch.Tensor, node_rows: Optional[torch.Tensor] = None):
        _require_cuda(row_ptr, "row_ptr")
        _require_cuda(col_idx, "col_idx")
        if row_ptr.dtype != torch.int64 or col_idx.dtype != torch.int32:
            raise ValueError("row_ptr must be int64 and col_idx int32")
        self._lib = N.load()
        self.row_ptr, self.col_idx, self.node_rows = row_ptr, col_idx, node_rows
        self.device = row_ptr.device
        self.V = int(row_ptr.numel()) - 1
        self.E = int(col_idx.numel())
        self._handle = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            if node_rows is None:
                N.check(self._lib.b200r_graph_create(self.device.index or 0, row_ptr.data_ptr(),
                                                     col_idx.data_ptr() if self.E else None, self.V, self.E, ctypes.byref(self._handle)))
            else:
                _require_cuda(node_rows, "node_rows")
                if node_rows.dtype != torch.int64 or node_rows.numel() != self.V:
                    raise ValueError("node_rows must be int64 [V]")
                N.check(self._lib.b200r_graph_create_mapped(self.device.index or 0, row_ptr.data_ptr(),
                                                            col_idx.data_ptr() if self.E else None, self.V, self.E,
                                                            node_rows.data_ptr(), ctypes.byref(self._handle)))

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.b200r_graph_destroy(self._handle)
            self._handle = ctypes.c_uint64(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        if not self._handle.value:
            raise RuntimeError("graph is closed")
        return self._handle.value


def expand_rescore(index: DenseIndex, graph: PassageGraph, Q: torch.Tensor, seeds: torch.Tensor, window: int,
                   alpha_dense: float, alpha_graph: float, k: int, check_status: bool = True, unbounded: bool = False) -> Dict[str, torch.Tensor]:
    """Graph second stage for B queries: BFS (<= window hops) from ``seeds`` [B, S] (i64, <0 = padding),
    cosine re-score, min-max fusion, top-k.  Returns device tensors keyed like the reference's meta keys.
    unbounded=True: queries whose neighbourhood exceeds the on-chip capacity are redone in global memory
    (b200r_expand_rescore_big) instead of being reported through out["status"]."""
    _require_cuda(Q, "Q")
    _require_cuda(seeds, "seeds")
    if Q.dtype != index.query_dtype or Q.dim() != 2 or Q.shape[1] != index.D:
        raise ValueError("Q must be [B, D] in the index's query dtype")
    if seeds.dtype != torch.int64 or seeds.dim() != 2 or seeds.shape[0] != Q.shape[0]:
        raise ValueError("seeds must be int64 [B, S]")
    B, S = int(seeds.shape[0]), int(seeds.shape[1])
    dev = index.device
    out = {
        "score": torch.empty((B, k), dtype=torch.float32, device=dev),
        "id": torch.empty((B, k), dtype=torch.int64, device=dev),
        "hop": torch.empty((B, k), dtype=torch.int32, device=dev),
        "score_graph_norm": torch.empty((B, k), dtype=torch.float32, device=dev),
        "score_dense_norm": torch.empty((B, k), dtype=torch.float32, device=dev),
        "cos": torch.empty((B, k), dtype=torch.float32, device=dev),
        "ncand": torch.empty((B,), dtype=torch.int32, device=dev),
        "status": torch.empty((B,), dtype=torch.int32, device=dev),
    }
    lib = N.load()
    # per-call scratch (parked candidate lists) from torch's caching allocator: stream-ordered, no sync
    ws = torch.empty(int(lib.b200r_expand_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(lib.b200r_expand_rescore(
            index.handle, graph.handle, Q.data_ptr(), seeds.data_ptr(), B, S, int(window), float(alpha_dense), float(alpha_graph), int(k),
            ws.data_ptr(), ws.numel(),
            out["score"].data_ptr(), out["id"].data_ptr(), out["hop"].data_ptr(), out["score_graph_norm"].data_ptr(),
            out["score_dense_norm"].data_ptr(), out["cos"].data_ptr(), out["ncand"].data_ptr(), out["status"].data_ptr(),
            _stream_ptr(dev)))
    if unbounded or check_status:
        over = (out["status"] != 0).nonzero().flatten().tolist()        # synchronises; almost always empty
        if over and unbounded:
            # the reference's BFS is unbounded (graph_utils.py:109-127): redo the few oversized queries in global memory
            big = torch.empty(int(lib.b200r_expand_big_workspace_bytes(graph.handle)), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                for q in over:
                    N.check(lib.b200r_expand_rescore_big(
                        index.handle, graph.handle, Q[q].data_ptr(), seeds[q].data_ptr(), S, int(window), float(alpha_dense),
                        float(alpha_graph), int(k), big.data_ptr(), big.numel(),
                        out["score"][q].data_ptr(), out["id"][q].data_ptr(), out["hop"][q].data_ptr(), out["score_graph_norm"][q].data_ptr(),
                        out["score_dense_norm"][q].data_ptr(), out["cos"][q].data_ptr(), out["ncand"][q:q + 1].data_ptr(), _stream_ptr(dev)))
                    out["status"][q] = 0
        elif over and check_status:
            raise N.B200RError(N.EOVERFLOW, f"{len(over)} queries reached more than {N.load().b200r_expand_max_candidates()} graph nodes")
    return out


def _expand_outputs(dev, B: int, k: int) -> Dict[str, torch.Tensor]:
    return {
        "score": torch.empty((B, k), dtype=torch.float32, device=dev),
        "id": torch.empty((B, k), dtype=torch.int64, device=dev),
        "hop": torch.empty((B, k), dtype=torch.int32, device=dev),
        "score_graph_norm": torch.empty((B, k), dtype=torch.float32, device=dev),
        "score_dense_norm": torch.empty((B, k), dtype=torch.float32, device=dev),
        "cos": torch.empty((B, k), dtype=torch.float32, device=dev),
        "ncand": torch.empty((B,), dtype=torch.int32, device=dev),
        "status": torch.empty((B,), dtype=torch.int32, device=dev),
    }


def _exchange_phase(index: DenseIndex, graph: "PassageGraph", comm, Q, seeds, window, alpha_dense, alpha_graph, k, n_total, phase, ws, out):
    B, S = int(seeds.shape[0]), int(seeds.shape[1])
    with torch.cuda.device(index.device):
        N.check(N.load().b200r_expand_rescore_exchange(
            index.handle, graph.handle, comm.handle, Q.data_ptr(), seeds.data_ptr(), B, S, int(window), float(alpha_dense),
            float(alpha_graph), int(k), int(n_total), int(phase), ws.data_ptr(), ws.numel(),
            out["score"].data_ptr(), out["id"].data_ptr(), out["hop"].data_ptr(), out["score_graph_norm"].data_ptr(),
            out["score_dense_norm"].data_ptr(), out["cos"].data_ptr(), out["ncand"].data_ptr(), out["status"].data_ptr(),
            _stream_ptr(index.device)))


def expand_rescore_exchange(index: DenseIndex, graph: "PassageGraph", comm, Q: torch.Tensor, seeds: torch.Tensor, window: int,
                            alpha_dense: float, alpha_graph: float, k: int, n_total_rows: int) -> Dict[str, torch.Tensor]:
    """Graph stage of ONE rank of a row-sharded corpus over the C-ABI exchange (one process per GPU): queries sharded by
    home rank, candidates by owner rank, every transfer a peer store from inside the kernels (b200r_expand_rescore_exchange).
    Every rank passes the same Q / seeds (global row or local node ids) and gets the same [B, k] results."""
    _require_cuda(Q, "Q")
    _require_cuda(seeds, "seeds")
    B = int(seeds.shape[0])
    out = _expand_outputs(index.device, B, k)
    ws = torch.empty(int(N.load().b200r_expand_workspace_bytes(B)), dtype=torch.uint8, device=index.device)
    _exchange_phase(index, graph, comm, Q, seeds, window, alpha_dense, alpha_graph, k, n_total_rows, 0, ws, out)
    return out


def expand_rescore_exchange_all(shards, graphs, comms, n_total_rows: int, Q_host: torch.Tensor, seeds_host: torch.Tensor, window: int,
                                alpha_dense: float, alpha_graph: float, k: int) -> Dict[str, torch.Tensor]:
    """The same stage with ONE host thread driving every rank (B200RetrievalBackend(devices=[...])): host Q / seeds are
    handed to every device, then phase p is issued on every rank before phase p+1 on any; rank 0's results are returned."""
    B = int(seeds_host.shape[0])
    Qd = [Q_host.to(s.device, non_blocking=True) for s in shards]
    if Q_host.dtype == torch.float32:            # provider embeddings: rounded to the query dtype on every device
        Qd = [s.to_query_dtype(q) for s, q in zip(shards, Qd)]
    Sd = [seeds_host.to(s.device, non_blocking=True) for s in shards]
    outs = [_expand_outputs(s.device, B, k) for s in shards]
    nbytes = int(N.load().b200r_expand_workspace_bytes(B))
    wss = [torch.empty(nbytes, dtype=torch.uint8, device=s.device) for s in shards]
    for phase in (1, 2, 3, 4):
        for g, s in enumerate(shards):
            _exchange_phase(s, graphs[g], comms[g], Qd[g], Sd[g], window, alpha_dense, alpha_graph, k, n_total_rows, phase, wss[g], outs[g])
    for s in shards:
        torch.cuda.synchronize(s.device)       # the windows / workspaces of all ranks stay alive until every rank is done
    return outs[0]


class ExpandShardState:
    """Device buffers of one rank's share of a sharded expand_rescore (see sharded.ShardedGraphStage).  The six [B, k]
    outputs are views of ONE packed record [id i64 | score f32 | hop i32 | graph_norm f32 | dense_norm f32 | cos f32]
    so that the ranks exchange them with a single all-gather."""

    FIELDS = (("id", torch.int64), ("score", torch.float32), ("hop", torch.int32), ("score_graph_norm", torch.float32),
              ("score_dense_norm", torch.float32), ("cos", torch.float32))

    def __init__(self, index: DenseIndex, B: int, k: int):
        dev = index.device
        lib = N.load()
        self.B, self.k = B, k
        self.n = (B * k + 1) // 2 * 2          # slots per field, even so that every field offset is 8-byte aligned
        self.ws = torch.empty(int(lib.b200r_expand_workspace_bytes(B)), dtype=torch.uint8, device=dev)
        self.minmax = torch.empty((B, 2), dtype=torch.float32, device=dev)
        self.record = torch.zeros(28 * self.n, dtype=torch.uint8, device=dev)
        self.out = self.views(self.record.unsqueeze(0), squeeze=True)
        self.out["ncand"] = torch.empty((B,), dtype=torch.int32, device=dev)
        self.out["status"] = torch.empty((B,), dtype=torch.int32, device=dev)

    def views(self, records: torch.Tensor, squeeze: bool = False) -> Dict[str, torch.Tensor]:
        """records: uint8 [G, 28*n] -> {field: [G, B, k] view} (shard stride 28*n bytes)."""
        G = int(records.shape[0])
        out, off = {}, 0
        for name, dt in self.FIELDS:
            size = torch.empty((), dtype=dt).element_size()
            v = records[:, off:off + size * self.n].view(dt)[:, :self.B * self.k].unflatten(1, (self.B, self.k))
            out[name] = v[0] if squeeze else v
            off += size * self.n
        return out


def expand_rescore_pass(index: DenseIndex, graph: PassageGraph, Q: torch.Tensor, seeds: torch.Tensor, window: int,
                        alpha_dense: float, alpha_graph: float, k: int, st: ExpandShardState, which: int,
                        own_lo: int, own_hi: int) -> None:
    """One pass (1 or 2) of the owner-computes graph stage on this rank's row shard; seeds/ids are GLOBAL row ids."""
    _require_cuda(Q, "Q")
    _require_cuda(seeds, "seeds")
    B, S = int(seeds.shape[0]), int(seeds.shape[1])
    o = st.out
    with torch.cuda.device(index.device):
        N.check(N.load().b200r_expand_rescore_sharded(
            index.handle, graph.handle, Q.data_ptr(), seeds.data_ptr(), B, S, int(window), float(alpha_dense), float(alpha_graph), int(k),
            int(which), int(own_lo), int(own_hi), st.ws.data_ptr(), st.ws.numel(), st.minmax.data_ptr(),
            o["score"].data_ptr(), o["id"].data_ptr(), o["hop"].data_ptr(), o["score_graph_norm"].data_ptr(),
            o["score_dense_norm"].data_ptr(), o["cos"].data_ptr(), o["ncand"].data_ptr(), o["status"].data_ptr(),
            _stream_ptr(index.device)))


def range_search(index: DenseIndex, Q: torch.Tensor, min_score: float, id_gt: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """EVERY row with score >= min_score for each query (no cap): -> (offsets i64 [B+1], ids i64 [total], scores f32 [total]);
    query q's hits are ids[offsets[q]:offsets[q+1]] in ascending id.  id_gt (i64 [B]): only ids > id_gt[q].  Two passes of the
    tile kernel (count, then fill into exactly-sized arrays, b200r_range_count / b200r_range_fill)."""
    _require_cuda(Q, "Q")
    if Q.dim() != 2 or Q.shape[1] != index.D or Q.dtype != index.query_dtype:
        raise ValueError(f"Q must be [B, {index.D}] {index.query_dtype}")
    if id_gt is not None:
        _require_cuda(id_gt, "id_gt")
        if id_gt.dtype != torch.int64 or id_gt.numel() != Q.shape[0]:
            raise ValueError("id_gt must be int64 [B]")
    lib = N.load()
    B, dev = int(Q.shape[0]), index.device
    nbytes = int(lib.b200r_range_workspace_bytes(index.handle, B))
    if nbytes == 0:
        raise N.B200RError(-1, N.last_error())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
    total = ctypes.c_int64(0)
    gt = id_gt.data_ptr() if id_gt is not None else None
    with torch.cuda.device(dev):
        N.check(lib.b200r_range_count(index.handle, Q.data_ptr(), B, float(min_score), gt, offsets.data_ptr(), ctypes.byref(total),
                                      ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
        n = int(total.value)
        ids = torch.empty(n, dtype=torch.int64, device=dev)
        scores = torch.empty(n, dtype=torch.float32, device=dev)
        N.check(lib.b200r_range_fill(index.handle, Q.data_ptr(), B, float(min_score), gt, n, ids.data_ptr(), scores.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    return offsets, ids, scores


def semantic_sim_edges(index: DenseIndex, threshold: float, batch: int = 4096, row_range: Optional[Tuple[int, int]] = None):
    """Corpus-scale version of EdgeBuilder.build step 4 (edge_builder.py:146-169): EVERY unordered pair (a, b), a < b, with
    cosine >= threshold (the reference enumerates itertools.combinations, i.e. each pair once, a before b) -- no per-row
    cap.  Rows [row_range) of the corpus are the `a` side (default: all), processed `batch` rows at a time through
    range_search with id_gt = the row's own id.  Yields (a_first, offsets i64 [n+1], dst i64 [m], sim f32 [m]) per block:
    row a_first + j pairs with dst[offsets[j]:offsets[j+1]] (ascending), all on the device."""
    lo, hi = row_range if row_range is not None else (0, index.N)
    for r0 in range(lo, hi, batch):
        r1 = min(hi, r0 + batch)
        Q = index.X[r0:r1]
        if Q.dtype != index.query_dtype or not Q.is_contiguous():
            Q = Q.to(index.query_dtype).contiguous()
        own = torch.arange(r0, r1, device=index.device, dtype=torch.int64) + index.id_base
        offsets, dst, sim = range_search(index, Q, threshold, id_gt=own)
        yield r0 + index.id_base, offsets, dst, sim
