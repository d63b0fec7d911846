"""B200RetrievalBackend -- drop-in for HybridRetrievalBackend behind the reference's retriever plug-in API.

Boundary mirrored (reference paths):
  * constructor keyword names / defaults      app/modules/retrieval/retrieval_backend.py:252-270
    (RetrievalAgentFlow.from_settings injects `router`/`sink` by NAME and silently drops unknown
    impl_kwargs, app/modules/retrieval/flow.py:92-108; RetrievalAdapter passes router/sink + kwargs,
    retrieval_adapter.py:37-42)
  * run(req) -> {"hits": [{"id","score","meta"}], "diagnostics": {...}}      retrieval_backend.py:303-385
  * retrieve(req) -> RetrievalOut                                             retrieval_backend.py:387-390
  * spans Backend/Expand|TextSearch|GraphExpand|DenseRerank                   retrieval_backend.py:308-332
  * Hit.meta keys read downstream (text, doc, sent_id, score_*_norm)         reasoning/strategies.py:108-117, 233-240
  * per-graph_id graphs: graph_root/<graph_id>/graph.json, q_match seeds      retrieval_backend.py:151-166, graph_utils.py:13-55

What changes underneath: instead of embedding <= bm25_pool_k BM25 candidates per query through the
provider and scoring them with a pure-Python cosine, the query embedding is searched EXACTLY against the
whole resident passage-embedding corpus on the GPU(s), the graph named by the request's `graph_id` (loaded
once, cached, never re-parsed per query) is expanded from its `q_match` seeds -- or from the dense top
`seed_k` passages when it has none -- and candidates are re-scored and fused on the device.  The optional
text channel stays on the host and is fused with the reference's formula.

One code path serves one request (`run` / `retrieve`) and many (`retrieve_many`, `run_batch`): queries are
embedded, searched, expanded and fused as arrays; Python objects are only built for the <= top_k winners.

The compute engine is injectable (tests drive the host logic with an oracle engine); the default engine is
the CUDA library -- one GPU, several GPUs driven by this process (`devices=[0, 1, ...]`, corpus row-sharded,
exchange inside libb200r) or one rank of a torchrun job -- and raises if it is unavailable: no CPU fallback.
"""
from __future__ import annotations

import collections
import contextlib
import gc
import itertools
import json
import logging
import os
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import store
from .dto import Hit, RetrievalOut

logger = logging.getLogger(__name__)


@contextlib.contextmanager
def _span(node: str, sink: Any, trace_id: str):
    """Use the host application's telemetry span when it is importable (app/telemetry/sinks.py:105-116)."""
    if sink is None:
        yield
        return
    try:
        from app.telemetry.sinks import span as host_span  # type: ignore
    except Exception:  # noqa: BLE001
        host_span = None
    if host_span is not None:
        with host_span(node, sink, trace_id):
            yield
        return
    import time
    t0 = time.time()
    sink.record({"trace_id": trace_id, "ts": t0, "event": "node_start", "node": node, "status": "running", "payload": {}})
    try:
        yield
        t1 = time.time()
        sink.record({"trace_id": trace_id, "ts": t1, "event": "node_end", "node": node, "status": "ok",
                     "duration_sec": t1 - t0, "payload": {}})
    except Exception as e:  # noqa: BLE001
        t1 = time.time()
        sink.record({"trace_id": trace_id, "ts": t1, "event": "error", "node": node, "status": "error",
                     "duration_sec": t1 - t0, "error": repr(e), "payload": {}})
        raise


GRAPH_KEYS = ("score", "id", "hop", "score_graph_norm", "score_dense_norm", "cos", "ncand", "status")


class CudaEngine:
    """Default engine: the corpus (row-sharded when several devices are given) and the graphs live in HBM; every call
    is a libb200r kernel sequence.  Arrays in, arrays out (numpy on the host side of the boundary)."""

    def __init__(self, emb_path: str, graph_path: str, device: str = "cuda:0", metric: str = "cos",
                 devices: Optional[Sequence[int]] = None, max_batch: int = 4096, max_k: int = 256):
        import torch
        from .sharded import shard_rows

        if not torch.cuda.is_available():
            raise RuntimeError("B200RetrievalBackend needs a CUDA device: there is no CPU fallback")
        devs = [torch.device("cuda", int(d)) for d in devices] if devices else [torch.device(device)]
        if max((d.index or 0) for d in devs) >= torch.cuda.device_count():
            raise RuntimeError(f"devices={list(devices or [device])} but only {torch.cuda.device_count()} CUDA devices are visible")
        X, meta, scales = store.load_embeddings(emb_path)
        tdtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32, "i8": torch.int8}[meta["dtype"]]
        N, D, G = int(meta["N"]), int(meta["D"]), len(devs)
        shards, row_scales = [], []
        for g, dev in enumerate(devs):
            lo, hi = shard_rows(N, G, g)
            host = torch.from_numpy(np.array(X[lo:hi], copy=True))      # the store is a read-only memmap
            if meta["dtype"] == "bf16":
                host = host.view(torch.bfloat16)
            if hi == lo:                                                    # an empty tail shard still needs a valid pointer
                host = torch.zeros((1, D), dtype=tdtype)
            shards.append(host.to(dev))
            row_scales.append(torch.from_numpy(np.array(scales[lo:hi], copy=True)).to(dev) if scales is not None else None)
        self._setup(shards, row_scales, N, metric or meta.get("metric", "cos"), max_batch, max_k)
        if graph_path:
            path = store.csr_path(graph_path)
            if not os.path.exists(path):
                raise FileNotFoundError(f"graph_path {graph_path!r} does not exist (store.save_csr writes '<name>.npz')")
            row_ptr, col_idx = store.load_csr(path)
            self.corpus_graphs = self._upload_graph(row_ptr, col_idx, None)

    @classmethod
    def from_shards(cls, shards: Sequence[Any], n_total: int, metric: str = "cos", max_batch: int = 4096, max_k: int = 256,
                    row_scales: Optional[Sequence[Any]] = None) -> "CudaEngine":
        """Engine over row shards that already live in HBM: shards[g] (a CUDA tensor on its device) holds rows
        shard_rows(n_total, G, g) of the corpus."""
        self = cls.__new__(cls)
        self._setup(list(shards), list(row_scales) if row_scales is not None else [None] * len(shards), int(n_total), metric, max_batch, max_k)
        return self

    def _setup(self, shards, row_scales, n_total: int, metric: str, max_batch: int, max_k: int) -> None:
        import torch
        from .index import DenseIndex, PassageGraph
        from .sharded import shard_rows
        self.torch = torch
        self.devs = [x.device for x in shards]
        self.device = self.devs[0]
        self.N, self.D = int(n_total), int(shards[0].shape[1])
        self.tdtype = shards[0].dtype
        G = len(shards)
        self._keep = list(zip(shards, row_scales))
        self.shards = []
        for g, (Xd, sc) in enumerate(self._keep):
            lo, _ = shard_rows(self.N, G, g)
            self.shards.append(DenseIndex(Xd, metric=metric, id_base=lo, row_scale=sc))
        self.query_dtype = self.shards[0].query_dtype
        self.max_batch, self.max_k = int(max_batch), int(max_k)
        self.comms = None
        if G > 1:
            from .comm import ShardComm
            self.comms = ShardComm.init_all([d.index for d in self.devs], self.max_batch, self.max_k, self.max_k)
        self._graph_cls = PassageGraph
        self.corpus_graphs = None       # one replica per device of the corpus-level CSR (graph_path)
        self._qpin = None               # pinned fp32 staging of the queries
        self._opin = None               # pinned staging of the results

    # -- helpers ------------------------------------------------------------------------------------------
    def _upload_graph(self, row_ptr, col_idx, node_rows):
        t = self.torch
        out = []
        for dev in self.devs:
            nr = t.from_numpy(np.ascontiguousarray(node_rows, dtype=np.int64)).to(dev) if node_rows is not None else None
            out.append(self._graph_cls(t.from_numpy(np.ascontiguousarray(row_ptr, dtype=np.int64)).to(dev),
                                       t.from_numpy(np.ascontiguousarray(col_idx, dtype=np.int32)).to(dev), nr))
        return out

    def _q_host(self, Q: np.ndarray):
        """f32 host queries -> PINNED fp32 staging (the H2D copies to the devices run asynchronously; the library rounds
        them to the index's query dtype on the device -- no 2-byte conversion on the host)."""
        q = self.torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32))
        n = q.numel()
        if self._qpin is None or self._qpin.numel() < n:
            self._qpin = self.torch.empty(max(n, 1), dtype=self.torch.float32).pin_memory()
        out = self._qpin[:n].view(q.shape)
        out.copy_(q)
        return out

    def _out_host(self, B: int, k: int):
        """Pinned result staging [B, k] (scores f32, ids i64): the D2H copies land here, then one memcpy into the caller's arrays."""
        n = B * k
        if self._opin is None or self._opin[0].numel() < n:
            self._opin = (self.torch.empty(max(n, 1), dtype=self.torch.float32).pin_memory(),
                          self.torch.empty(max(n, 1), dtype=self.torch.int64).pin_memory())
        return self._opin[0][:n].view(B, k), self._opin[1][:n].view(B, k)

    def has_graph(self) -> bool:
        return self.corpus_graphs is not None

    # -- dense search ---------------------------------------------------------------------------------------
    def dense_search(self, Q: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Q f32 [B, D] -> (scores f32 [B, k], ids i64 [B, k]); batches beyond max_batch run in slices."""
        import ctypes
        from . import _native as N
        B = int(Q.shape[0])
        scores = np.empty((B, k), dtype=np.float32)
        ids = np.empty((B, k), dtype=np.int64)
        if len(self.devs) > 1 and k > self.max_k:
            raise ValueError(f"k={k} exceeds the exchange window (max_k={self.max_k}); construct the engine with a larger max_k")
        for b0 in range(0, B, self.max_batch):
            b1 = min(B, b0 + self.max_batch)
            qh = self._q_host(Q[b0:b1])
            hs, hi = self._out_host(b1 - b0, k)
            if len(self.devs) == 1:
                self.shards[0].search_host(qh, k, out=(hs, hi))
            else:
                G = len(self.devs)
                ih = (ctypes.c_uint64 * G)(*[s.handle for s in self.shards])
                ch = (ctypes.c_uint64 * G)(*[c.handle for c in self.comms])
                N.check(N.load().b200r_search_sharded_host_all_f32(ih, ch, G, qh.data_ptr(), b1 - b0, k, N.PATH_AUTO,
                                                                   hs.data_ptr(), hi.data_ptr(), None))
            scores[b0:b1], ids[b0:b1] = hs.numpy(), hi.numpy()
        return scores, ids

    # -- MMR post-filter ---------------------------------------------------------------------------------------
    def mmr(self, ids: np.ndarray, scores: np.ndarray, top_k: int, lam: float) -> np.ndarray:
        """[B, P] candidate rows / scores -> positions [B, top_k] of the MMR picks (app/utils/similarity.py:33-62)."""
        from .index import mmr_diversify
        if len(self.devs) > 1:
            raise NotImplementedError("MMR needs the candidates' rows on one device; it is not available with devices=[...]")
        t = self.torch
        _, _, pos = mmr_diversify(self.shards[0], t.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(self.device),
                                  t.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).to(self.device), top_k, lam)
        return pos.cpu().numpy()

    # -- graph stage ------------------------------------------------------------------------------------------
    def expand_rescore(self, Q: np.ndarray, seeds: np.ndarray, window: int, alpha_dense: float, alpha_graph: float, k: int,
                       graph: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """graph = None: the corpus-level CSR (node id == row); else (row_ptr, col_idx, node_rows) of a mapped graph
        (local node ids, e.g. the block-diagonal union of the batch's per-graph_id graphs; seeds are local ids)."""
        from .index import expand_rescore, expand_rescore_exchange_all
        t = self.torch
        graphs = self.corpus_graphs if graph is None else self._upload_graph(*graph)
        if graphs is None:
            raise RuntimeError("no graph loaded")
        out: Dict[str, List[np.ndarray]] = {key: [] for key in GRAPH_KEYS}
        B = int(Q.shape[0])
        for b0 in range(0, B, self.max_batch):
            b1 = min(B, b0 + self.max_batch)
            qh = self._q_host(Q[b0:b1])
            sh = t.from_numpy(np.ascontiguousarray(seeds[b0:b1], dtype=np.int64))
            if len(self.devs) == 1:
                qd = self.shards[0].to_query_dtype(qh.to(self.device, non_blocking=True))
                res = expand_rescore(self.shards[0], graphs[0], qd, sh.to(self.device), window,
                                     alpha_dense, alpha_graph, k, check_status=False, unbounded=True)
            else:
                res = expand_rescore_exchange_all(self.shards, graphs, self.comms, self.N, qh, sh, window, alpha_dense, alpha_graph, k)
            for key in GRAPH_KEYS:
                out[key].append(res[key].cpu().numpy())
        return {key: np.concatenate(v) for key, v in out.items()}


class LocalGraph:
    """One graph_id's graph.json reduced to what the hot path needs (build_index, graph_utils.py:25-55): local CSR of the
    undirected next_in_doc adjacency, node -> corpus row, q_match seeds, and the sentence texts of the lexical seed fallback."""

    __slots__ = ("row_ptr", "col_idx", "node_rows", "qmatch", "tokens", "row_to_node")

    def __init__(self, row_ptr, col_idx, node_rows, qmatch, tokens):
        self.row_ptr, self.col_idx, self.node_rows, self.qmatch, self.tokens = row_ptr, col_idx, node_rows, qmatch, tokens
        self.row_to_node = {int(r): i for i, r in enumerate(node_rows)}

    @property
    def V(self) -> int:
        return len(self.node_rows)


def _tokenize(text: str) -> List[str]:
    """graph_utils.py:9-10."""
    return [t for t in re.split(r"[^a-zA-Z0-9]+", (text or "").lower()) if t]


@dataclass
class B200RetrievalBackend:
    router: Any = None
    sink: Any = None

    # --- names shared with HybridRetrievalBackend (retrieval_backend.py:256-270) ---
    index_path: str = "data/hotpotqa/docs.jsonl"
    bm25_k1: float = 1.5
    bm25_b: float = 0.75
    graph_root: str = "data/graph"
    graph_window: int = 2
    alpha_text: float = 0.4
    alpha_graph: float = 0.2
    alpha_dense: float = 0.4
    bm25_pool_k: int = 150
    default_top_k: int = 20
    qe_lines: int = 3
    qe_attr_paraphrase: bool = True
    embed_batch: int = 50

    # --- new: where the resident corpus / graph come from and how they are searched ---
    emb_path: str = "data/hotpotqa/embeddings"      # store.save_embeddings directory, rows aligned with index_path
    graph_path: str = ""                            # optional corpus-level next_in_doc CSR (store.save_csr), used for a
                                                    # graph_id that has no graph_root/<graph_id>/graph.json
    seed_k: int = 20                                # dense top-S passages that seed the expansion when a graph has no q_match
    device: str = "cuda:0"
    devices: Optional[List[int]] = None             # e.g. [0,...,7]: this process drives all of them, corpus row-sharded
    metric: str = "cos"
    graph_cache_size: int = 4096                    # per-graph_id graphs kept parsed (LRU)
    mmr_lambda: Optional[float] = None              # not None: the final top_k are picked from the best `mmr_pool` fused candidates by
    mmr_pool: int = 100                             # MMR on the device (mmr_diversify, app/utils/similarity.py:33-62; item vectors = corpus rows)
    pause_gc: bool = True                           # retrieve_many: cyclic GC off while a batch's results are materialised
    text_searcher: Any = None                       # optional object with .search(queries=, top_k=) -> raw hits (host BM25)
    expander: Any = None                            # optional object with .expand(query=, trace_id=) -> List[str]
    engine: Any = None                              # injectable compute engine (tests); default CudaEngine

    docs: List[Dict[str, Any]] = field(default_factory=list, repr=False)

    def __post_init__(self):
        if not self.docs:
            self.docs = store.load_docs(self.index_path)
        if self.engine is None:
            self.engine = CudaEngine(self.emb_path, self.graph_path, self.device, self.metric, devices=self.devices)
        n = getattr(self.engine, "N", None)
        if self.docs and n is not None and n != len(self.docs):
            raise ValueError(f"embedding rows ({n}) and docs.jsonl lines ({len(self.docs)}) disagree")
        self._graphs: "collections.OrderedDict[str, Optional[LocalGraph]]" = collections.OrderedDict()
        self._node_rows: Optional[Dict[str, int]] = None
        logger.info("[B200RetrievalBackend] init: rows=%s seed_k=%d alpha_text=%s alpha_graph=%s alpha_dense=%s graph_window=%d engine=%s",
                    n, self.seed_k, self.alpha_text, self.alpha_graph, self.alpha_dense, self.graph_window, type(self.engine).__name__)

    # ------------------------------------------------------------------------------------------ helpers
    def _resolve_embed_model(self) -> str:
        """Same resolution order as DenseReranker._resolve_embed_model (retrieval_backend.py:199-213)."""
        try:
            policy = getattr(self.router, "policy", {}) or {}
            emb = policy.get("embedding") or []
            if emb and isinstance(emb[0], dict) and emb[0].get("model"):
                return emb[0]["model"]
            provider_name = policy.get("embedding_provider")
            providers = getattr(self.router, "providers", {}) or {}
            prov = providers.get(provider_name) if provider_name else None
            if prov and hasattr(prov, "kwargs"):
                m = (getattr(prov, "kwargs") or {}).get("embed_model")
                if m:
                    return m
        except Exception:  # noqa: BLE001
            pass
        return "text-embedding-3-large"

    def _embed_queries(self, queries: Sequence[str], trace_ids: Sequence[str]) -> List[Optional[List[float]]]:
        """router.embed as DenseReranker.score calls it (retrieval_backend.py:227-231: dict or list result, failures
        degrade to "no dense channel"); a batch is embedded in chunks of max(8, embed_batch) texts like :234-243."""
        model = self._resolve_embed_model()
        out: List[Optional[List[float]]] = [None] * len(queries)
        if len(queries) == 1:
            chunks = [(0, 1)]
        else:
            step = max(8, int(self.embed_batch))
            chunks = [(i, min(len(queries), i + step)) for i in range(0, len(queries), step)]
        for a, b in chunks:
            try:
                emb = self.router.embed(model_hint=model, texts=list(queries[a:b]), require={"trace_id": trace_ids[a]})
                vecs = emb.get("vectors") if isinstance(emb, dict) else emb
                for j in range(a, b):
                    out[j] = vecs[j - a]
            except Exception as e:  # noqa: BLE001
                logger.error("[B200RetrievalBackend] query embed error: %s", e, exc_info=True)
        return out

    @staticmethod
    def _query_matrix(qvs: Sequence[Any], dim: Optional[int]) -> Tuple[List[int], Optional[np.ndarray]]:
        """Provider vectors (lists of Python floats or arrays) -> (requests with a usable embedding, f32 [len(valid), D]).
        Unusable = missing, of the wrong width, or all zero (the reference's _cosine would score 0.0 everywhere)."""
        cand = []
        for i, qv in enumerate(qvs):
            if qv is None:
                continue
            if len(qv) == 0 or (dim is not None and len(qv) != dim) or (cand and len(qv) != len(qvs[cand[0]])):
                logger.warning("[B200RetrievalBackend] query embedding unusable (dim=%s, corpus dim=%s): dense channel empty", len(qv), dim)
                continue
            cand.append(i)
        if not cand:
            return [], None
        D = len(qvs[cand[0]])
        if all(isinstance(qvs[i], np.ndarray) for i in cand):
            Q = np.stack([qvs[i] for i in cand]).astype(np.float32, copy=False)
        else:       # one flat pass over the Python floats: ~1.5x faster than np.asarray on a list of lists
            Q = np.fromiter(itertools.chain.from_iterable(qvs[i] for i in cand), dtype=np.float32, count=len(cand) * D).reshape(len(cand), D)
        nz = np.any(Q != 0, axis=1)
        if not nz.all():
            logger.warning("[B200RetrievalBackend] %d all-zero query embeddings: dense channel empty for them", int((~nz).sum()))
            cand = [i for i, ok in zip(cand, nz.tolist()) if ok]
            Q = np.ascontiguousarray(Q[nz])
        return cand, (Q if cand else None)

    def _row_meta(self, row: int) -> Dict[str, Any]:
        d = self.docs[row] if 0 <= row < len(self.docs) else {}
        return {"kind": "sentence", "text": d.get("text"), "doc": d.get("title"), "sent_id": d.get("sent_id"), "row": row}

    @staticmethod
    def _normalize_id(hit: Dict[str, Any]) -> Tuple[str, Dict[str, Any]]:
        """Literal mirror of HybridRetrievalBackend._normalize_id (retrieval_backend.py:283-294), including its
        treatment of a falsy sent_id (sentence 0 yields 'sent::<doc>::')."""
        _id = hit.get("id") or ""
        meta = hit.get("meta") or {}
        doc = meta.get("doc") or meta.get("title")
        sid = meta.get("sent_id") or meta.get("sid")
        if doc is not None and sid is not None:
            norm = f"sent::{doc}::{sid}"
        elif doc is not None:
            norm = f"sent::{doc}::{sid or ''}"
        else:
            norm = _id or "sent::unknown::"
        return norm, meta

    @staticmethod
    def _minmax_norm(values: Dict[str, float]) -> Dict[str, float]:
        """retrieval_backend.py:296-301."""
        if not values:
            return {}
        vs = list(values.values())
        vmin, vmax = min(vs), max(vs)
        if vmax <= vmin:
            return {k: 0.0 for k in values}
        return {k: (v - vmin) / (vmax - vmin) for k, v in values.items()}

    # ------------------------------------------------------------------------------------------ per-graph_id graphs
    def _node_to_row(self, node_id: str) -> Optional[int]:
        """Sentence node ids are '<title>::sent<j>' (NodeBuilder.build, graph_construction/node_builder.py:61-66); the corpus
        row is the docs.jsonl line with that title and sent_id (first one when a paragraph appears under several questions)."""
        if self._node_rows is None:
            table: Dict[str, int] = {}
            for row, d in enumerate(self.docs):
                table.setdefault(f"{d.get('title')}::sent{d.get('sent_id')}", row)
            self._node_rows = table
        return self._node_rows.get(node_id)

    def graph_for(self, graph_id: str) -> Optional[LocalGraph]:
        """load_graph_json + build_index (graph_utils.py:13-55) ONCE per graph_id (LRU), instead of per query
        (retrieval_backend.py:154-155).  None when graph_root/<graph_id>/graph.json does not exist."""
        if not graph_id:
            return None
        if graph_id in self._graphs:
            self._graphs.move_to_end(graph_id)
            return self._graphs[graph_id]
        path = os.path.join(self.graph_root or "", graph_id, "graph.json")
        lg: Optional[LocalGraph] = None
        if self.graph_root and os.path.exists(path):
            with open(path, "r", encoding="utf-8") as f:
                graph = json.load(f)
            row_ptr, col_idx, node_rows, qmatch, texts = store.graph_json_to_local(graph, self._node_to_row)
            lg = LocalGraph(row_ptr, col_idx, node_rows, qmatch, [set(_tokenize(t)) if t else None for t in texts])
        self._graphs[graph_id] = lg
        while len(self._graphs) > max(1, int(self.graph_cache_size)):
            self._graphs.popitem(last=False)
        return lg

    def _graph_seeds(self, lg: LocalGraph, query: str, dense_rows: Sequence[int]) -> List[int]:
        """Seeds of expand_qmatch_neighbors (graph_utils.py:73-83): the q_match targets; else every sentence node whose
        text shares a token with the query; else (no counterpart in the reference, whose result would be empty) the dense
        top seed_k passages that are nodes of this graph."""
        if lg.qmatch:
            return list(lg.qmatch)
        q_terms = set(_tokenize(query))
        seeds = [i for i, toks in enumerate(lg.tokens) if toks and (q_terms & toks)]
        if seeds:
            return seeds
        return [lg.row_to_node[r] for r in dense_rows[: self.seed_k] if r in lg.row_to_node]

    # ------------------------------------------------------------------------------------------ the call
    def run_batch(self, reqs: Sequence[Any]) -> "RetrievalBatch":
        """Embeds, searches, expands and fuses every request of the batch as arrays (ONE dense search, ONE expansion per
        distinct window); returns a columnar RetrievalBatch whose rows are materialised into hits on demand."""
        n = len(reqs)
        trace_ids = [getattr(r, "trace_id", None) or "trace-demo" for r in reqs]
        top_ks = [int(getattr(r, "top_k", None) or self.default_top_k) for r in reqs]
        t0 = trace_ids[0] if n else "trace-demo"
        with _span("Backend/Expand", self.sink, t0):
            queries_per_req = [[r.query] + (list(self.expander.expand(query=r.query, trace_id=t)) if self.expander is not None else [])
                               for r, t in zip(reqs, trace_ids)]
        with _span("Backend/TextSearch", self.sink, t0):
            text_hits = [list(self.text_searcher.search(queries=qs, top_k=max(k, self.bm25_pool_k))) if self.text_searcher is not None else []
                         for qs, k in zip(queries_per_req, top_ks)]
        windows = []
        for r in reqs:
            gw = getattr(r, "graph_window", None)
            windows.append(max(0, int(gw if isinstance(gw, int) else self.graph_window)))

        batch = RetrievalBatch(self, reqs, trace_ids, top_ks, queries_per_req, text_hits, windows)
        dim = getattr(self.engine, "D", None)
        with _span("Backend/DenseRerank", self.sink, t0):
            qvs = self._embed_queries([r.query for r in reqs], trace_ids)
            valid, Q = self._query_matrix(qvs, dim)
            if valid:
                pool = max(self.seed_k, max(top_ks[i] for i in valid))
                s, ids = self.engine.dense_search(Q, pool)
                batch.set_dense(valid, Q, np.asarray(s, dtype=np.float32), np.asarray(ids, dtype=np.int64))

        with _span("Backend/GraphExpand", self.sink, t0):
            if valid and self.alpha_graph != 0.0:
                self._graph_stage(batch, valid, Q)
        if self.mmr_lambda is not None and valid:
            batch.apply_mmr(float(self.mmr_lambda), int(self.mmr_pool))
        return batch

    def _graph_stage(self, batch: "RetrievalBatch", valid: List[int], Q: np.ndarray) -> None:
        """Group the requests that take the graph channel by (kind of graph, window) and expand each group in one call.
        Requests whose graph_id has its own graph.json are expanded together over the block-diagonal union of their
        graphs (local node ids, shared blocks for repeated graph_ids); the others use the corpus-level CSR, if any."""
        reqs = batch.reqs
        kk = max(max(batch.top_ks[i] for i in valid), self.bm25_pool_k)
        groups: Dict[Tuple[str, int], List[int]] = {}
        local: Dict[int, LocalGraph] = {}
        has_dense = (batch._dense_i[:, 0] >= 0)
        for pos, i in enumerate(valid):
            gid = getattr(reqs[i], "graph_id", "") or ""
            if not gid or not has_dense[pos]:
                continue
            lg = self.graph_for(gid)
            if lg is not None:
                if lg.V == 0:
                    continue                                  # an empty graph.json: no graph channel, like the reference
                local[i] = lg
                groups.setdefault(("local", batch.windows[i]), []).append(pos)
            elif self.engine.has_graph():
                groups.setdefault(("corpus", batch.windows[i]), []).append(pos)
        for (kind, window), members in groups.items():
            idx = [valid[p] for p in members]
            Qg = Q[members]
            if kind == "corpus":
                # seeds = the dense top-seed_k rows (ids are already padded with -1 where the corpus has fewer rows)
                seeds = np.ascontiguousarray(batch._dense_i[members, : self.seed_k])
                out = self.engine.expand_rescore(Qg, seeds, window, self.alpha_dense, self.alpha_graph, kk)
            else:
                blocks: Dict[int, int] = {}                     # id(LocalGraph) -> node offset of its block
                rp, ci, nr = [np.zeros(1, dtype=np.int64)], [], []
                node_off = edge_off = 0
                seed_lists = []
                for i in idx:
                    lg = local[i]
                    if id(lg) not in blocks:
                        blocks[id(lg)] = node_off
                        rp.append(lg.row_ptr[1:] + edge_off)
                        ci.append(lg.col_idx.astype(np.int64) + node_off)
                        nr.append(lg.node_rows)
                        node_off += lg.V
                        edge_off += int(lg.row_ptr[-1])
                    base = blocks[id(lg)]
                    seed_lists.append([base + s for s in self._graph_seeds(lg, reqs[i].query, batch.dense_rows(i).tolist())])
                S = max(1, max(len(sl) for sl in seed_lists))
                seeds = np.full((len(idx), S), -1, dtype=np.int64)
                for j, sl in enumerate(seed_lists):
                    seeds[j, : len(sl)] = sl
                graph = (np.concatenate(rp), np.concatenate(ci).astype(np.int32) if ci else np.zeros(0, np.int32), np.concatenate(nr))
                out = self.engine.expand_rescore(Qg, seeds, window, self.alpha_dense, self.alpha_graph, kk, graph=graph)
            batch.set_graph(idx, {key: np.asarray(out[key]) for key in GRAPH_KEYS})

    def run(self, req: Any) -> Dict[str, Any]:
        return self.run_batch([req]).result(0)

    def retrieve(self, req: Any) -> RetrievalOut:
        return self.run_batch([req]).output(0)

    def retrieve_many(self, reqs: Sequence[Any]) -> List[RetrievalOut]:
        """Batching shim: the orchestrator issues one query at a time
        (app/orchestrator/nodes.py:110-115, my_code/run_system.py:35-39); a batch runner hands all requests over at once
        so the dense stage is ONE search launch and the graph stage ONE expand launch.  Identical results to retrieve()
        per request (it is the same code path with a batch of one)."""
        if not reqs:
            return []
        batch = self.run_batch(reqs)
        # Thousands of short-lived, acyclic dicts / DTOs per call: with the collector on, every few hundred requests trigger
        # a generation-2 pass over the resident docs table (hundreds of thousands of dicts) -- a quarter of the call's time.
        pause = self.pause_gc and len(reqs) > 1 and gc.isenabled()
        if pause:
            gc.disable()
        try:
            return [batch.output(i) for i in range(len(reqs))]
        finally:
            if pause:
                gc.enable()


class RetrievalBatch:
    """Columnar result of B200RetrievalBackend.run_batch: per request the candidate rows, fused scores and channel
    scores as arrays.  `result(i)` / `output(i)` materialise request i in the reference's shapes
    (retrieval_backend.py:365-385); batch runners that only need rows and scores read `rows(i)` / `scores(i)` and never
    pay for Python objects."""

    def __init__(self, be: B200RetrievalBackend, reqs, trace_ids, top_ks, queries, text_hits, windows):
        self.be, self.reqs, self.trace_ids, self.top_ks = be, list(reqs), trace_ids, top_ks
        self.queries, self.text_hits, self.windows = queries, text_hits, windows
        self._pos: Dict[int, int] = {}              # request -> row of the dense arrays
        self._dense_s: Optional[np.ndarray] = None
        self._dense_i: Optional[np.ndarray] = None
        self._graph: Dict[int, Tuple[Dict[str, np.ndarray], int]] = {}   # request -> (group arrays, row in them)
        self._order: Dict[int, np.ndarray] = {}     # request -> MMR order of its candidates
        self._dense_dn: Optional[np.ndarray] = None
        self._heads: Dict[int, Tuple[int, Dict[str, list]]] = {}   # id(group arrays) -> (columns converted, Python lists of them)
        self._embed_model: Optional[str] = None

    # -- filled by the backend ------------------------------------------------------------------------------
    def set_dense(self, valid: List[int], Q: np.ndarray, scores: np.ndarray, ids: np.ndarray) -> None:
        self._pos = {i: p for p, i in enumerate(valid)}
        self._dense_s, self._dense_i = scores, ids

    def _dense_norm(self) -> np.ndarray:
        """Min-max of every request's dense pool (retrieval_backend.py:296-301) for the whole batch at once: f64 [n, pool]."""
        if self._dense_dn is None:
            sc = self._dense_s.astype(np.float64)
            ok = self._dense_i >= 0
            lo = np.where(ok, sc, np.inf).min(axis=1, keepdims=True)
            hi = np.where(ok, sc, -np.inf).max(axis=1, keepdims=True)
            with np.errstate(invalid="ignore", divide="ignore"):
                dn = (sc - lo) / (hi - lo)
            self._dense_dn = np.where(hi > lo, dn, 0.0)
        return self._dense_dn

    def set_graph(self, idx: List[int], out: Dict[str, np.ndarray]) -> None:
        for j, i in enumerate(idx):
            self._graph[i] = (out, j)

    def apply_mmr(self, lam: float, pool: int) -> None:
        """Re-order every request's best `pool` candidates by MMR (ONE device call for the batch); the rest keep their order."""
        idx = [i for i in range(len(self.reqs)) if len(self.columns(i)["rows"])]
        if not idx:
            return
        P = max(1, min(pool, max(len(self.columns(i)["rows"]) for i in idx)))
        ids = np.full((len(idx), P), -1, dtype=np.int64)
        sc = np.full((len(idx), P), -np.inf, dtype=np.float32)
        for j, i in enumerate(idx):
            c = self.columns(i)
            n = min(P, len(c["rows"]))
            ids[j, :n], sc[j, :n] = c["rows"][:n], c["fused"][:n]
        pos = np.asarray(self.be.engine.mmr(ids, sc, P, lam))
        for j, i in enumerate(idx):
            n = len(self.columns(i)["rows"])
            head = [int(x) for x in pos[j] if x >= 0]
            self._order[i] = np.asarray(head + list(range(min(P, n), n)), dtype=np.int64)

    def dense_rows(self, i: int) -> np.ndarray:
        p = self._pos.get(i)
        if p is None:
            return np.zeros(0, dtype=np.int64)
        ids = self._dense_i[p]
        return ids[ids >= 0]

    # -- columnar accessors ---------------------------------------------------------------------------------
    def columns(self, i: int) -> Dict[str, np.ndarray]:
        """Candidates of request i BEFORE the text channel / top_k cut, best first (MMR order if enabled): rows, fused, gnorm,
        dnorm, hop, cos."""
        c = self._columns(i)
        order = self._order.get(i)
        if order is not None:
            c = {k: (v[order] if isinstance(v, np.ndarray) else v) for k, v in c.items()}
            c["mmr"] = True
        return c

    def _columns(self, i: int) -> Dict[str, np.ndarray]:
        if i in self._graph:
            out, j = self._graph[i]
            keep = out["id"][j] >= 0
            return {"rows": out["id"][j][keep], "fused": out["score"][j][keep].astype(np.float64),
                    "gnorm": out["score_graph_norm"][j][keep], "dnorm": out["score_dense_norm"][j][keep],
                    "hop": out["hop"][j][keep], "cos": out["cos"][j][keep], "ncand": int(out["ncand"][j]),
                    "status": int(out["status"][j]), "graph": True}
        p = self._pos.get(i)
        if p is None:
            z = np.zeros(0)
            return {"rows": np.zeros(0, np.int64), "fused": z, "gnorm": z, "dnorm": z, "hop": np.zeros(0, np.int32), "cos": z,
                    "ncand": 0, "status": 0, "graph": False}
        keep = self._dense_i[p] >= 0
        rows, sc = self._dense_i[p][keep], self._dense_s[p][keep].astype(np.float64)
        # dense-only: the candidate set is the exact dense top-k; min-max over it like the reference does over its pool (:296-301)
        dn = self._dense_norm()[p][keep]
        return {"rows": rows, "fused": self.be.alpha_dense * dn, "gnorm": np.zeros_like(sc), "dnorm": dn,
                "hop": np.full(len(sc), -1, np.int32), "cos": sc, "ncand": 0, "status": 0, "graph": False}

    def rows(self, i: int) -> np.ndarray:
        return self.columns(i)["rows"][: self.top_ks[i]]

    def scores(self, i: int) -> np.ndarray:
        return self.columns(i)["fused"][: self.top_ks[i]]

    # -- the reference's shapes -----------------------------------------------------------------------------
    def _head(self, i: int, need: int):
        """Request i's leading candidates as Python lists, converted for its whole group in ONE tolist() per column (numpy
        scalar reads and per-request slicing are what the object path otherwise spends its time on).  Returns
        (lists, row in them, is_graph, all columns converted) or None when the request has no candidates."""
        if i in self._graph:
            out, j = self._graph[i]
            src, width = out, out["id"].shape[1]
        else:
            j = self._pos.get(i)
            if j is None:
                return None
            src, width = None, self._dense_i.shape[1]
        key = id(src) if src is not None else 0
        have = self._heads.get(key)
        if have is None or (have[0] < need and have[0] < width):
            m = min(width, max(need, have[0] * 2 if have else 0))
            if src is not None:
                lists = {"rows": src["id"][:, :m].tolist(), "fused": src["score"][:, :m].tolist(), "gnorm": src["score_graph_norm"][:, :m].tolist(),
                         "dnorm": src["score_dense_norm"][:, :m].tolist(), "hop": src["hop"][:, :m].tolist(), "cos": src["cos"][:, :m].tolist(),
                         "ncand": src["ncand"].tolist(), "status": src["status"].tolist()}
            else:
                dn = self._dense_norm()[:, :m]
                lists = {"rows": self._dense_i[:, :m].tolist(), "fused": (self.be.alpha_dense * dn).tolist(), "dnorm": dn.tolist(),
                         "cos": self._dense_s[:, :m].tolist()}
            have = (m, lists)
            self._heads[key] = have
        return have[1], j, src is not None, have[0] >= width

    def _fast_hits(self, i: int, top_k: int):
        """result()'s hit list for a request without a text channel or MMR order, from the batch-converted lists; None when
        the converted columns do not cover it (the general path then runs)."""
        head = self._head(i, top_k + 8)
        if head is None:
            return [], 0
        L, j, is_graph, complete = head
        if is_graph and L["status"][j] != 0:
            return None
        rows, fs, cs, dns = L["rows"][j], L["fused"][j], L["cos"][j], L["dnorm"][j]
        gns = L["gnorm"][j] if is_graph else None
        hops = L["hop"][j] if is_graph else None
        docs, ndocs = self.be.docs, len(self.be.docs)
        # the general path looks at the first min(valid candidates, top_k + 8) rows and stops at top_k distinct ids
        fused: Dict[str, Dict[str, Any]] = {}
        taken = 0
        exhausted = True
        for idx, r in enumerate(rows):
            if r < 0:
                continue
            if taken == top_k + 8:
                exhausted = False
                break
            taken += 1
            d = docs[r] if r < ndocs else {}
            hop = hops[idx] if is_graph else -1
            title, sid = d.get("title"), d.get("sent_id")
            meta = {"kind": "sentence", "text": d.get("text"), "doc": title, "sent_id": sid, "row": r,
                    "source": "graph" if hop > 0 else "dense"}
            if hop >= 0:
                meta["distance"] = hop
            meta["score_dense"] = cs[idx]
            meta["score_text_norm"] = 0.0
            meta["score_graph_norm"] = gns[idx] if is_graph else 0.0
            meta["score_dense_norm"] = dns[idx]
            # _normalize_id on this meta (retrieval_backend.py:283-294): a falsy title / sent_id counts as missing
            nid = (f"sent::{title}::{sid}" if sid else f"sent::{title}::") if title else f"sent::row{r}"
            sc = fs[idx]
            prev = fused.get(nid)
            if prev is None or sc > prev["score"]:
                fused[nid] = {"id": nid, "score": sc, "meta": meta, "_row": r}
            if len(fused) >= top_k:
                exhausted = False
                break
        if exhausted and not complete:
            return None                                # ran off the converted columns before the general path would stop
        ordered = sorted(fused.values(), key=lambda h: h["_row"])
        ordered.sort(key=lambda h: h["score"], reverse=True)
        ncand = L["ncand"][j] if is_graph else 0
        return [{"id": h["id"], "score": h["score"], "meta": h["meta"]} for h in ordered[:top_k]], ncand

    def _diagnostics(self, i: int, top_k: int, n_text: int, ncand: int, n_rows: int) -> Dict[str, Any]:
        be = self.be
        p = self._pos.get(i)
        has_dense = p is not None and self._dense_i.shape[1] > 0 and bool(self._dense_i[p, 0] >= 0)
        if self._embed_model is None:
            self._embed_model = be._resolve_embed_model()
        return {
            "queries": self.queries[i],
            "bm25_candidates": n_text,
            "graph_candidates": ncand,
            "dense_scored": getattr(be.engine, "N", n_rows) if has_dense else 0,
            "weights": {"alpha_text": be.alpha_text, "alpha_graph": be.alpha_graph, "alpha_dense": be.alpha_dense},
            "pool": {"bm25_pool_k": be.bm25_pool_k, "final_top_k": top_k, "seed_k": be.seed_k},
            "graph_window_used": self.windows[i],
            "embed_batch": int(be.embed_batch),
            "resolved_embed_model": self._embed_model,
            "engine": type(be.engine).__name__,
        }

    def result(self, i: int) -> Dict[str, Any]:
        be, top_k = self.be, self.top_ks[i]
        if not (self.text_hits[i] and be.alpha_text != 0.0) and i not in self._order:
            fast = self._fast_hits(i, top_k)
            if fast is not None:
                hits, ncand = fast
                return {"hits": hits, "diagnostics": self._diagnostics(i, top_k, len(self.text_hits[i]), ncand, len(hits))}
        c = self.columns(i)
        if c["status"] != 0:
            raise RuntimeError(f"graph expansion of request {i} reached more nodes than the expansion kernels hold "
                               f"({c['ncand']}); lower graph_window")
        t_hits_raw = self.text_hits[i]
        use_text = bool(t_hits_raw) and be.alpha_text != 0.0
        n_take = len(c["rows"]) if use_text else min(len(c["rows"]), top_k + 8)   # a few spare rows for id collisions
        fused: Dict[str, Dict[str, Any]] = {}
        for j in range(n_take):
            r = int(c["rows"][j])
            meta = be._row_meta(r)
            hop = int(c["hop"][j])
            meta["source"] = "graph" if (c["graph"] and hop > 0) else "dense"
            if hop >= 0:
                meta["distance"] = hop
            meta["score_dense"] = float(c["cos"][j])
            nid, _ = be._normalize_id({"id": f"sent::row{r}", "meta": meta})
            meta["score_text_norm"], meta["score_graph_norm"], meta["score_dense_norm"] = 0.0, float(c["gnorm"][j]), float(c["dnorm"][j])
            sc = float(c["fused"][j])
            if nid not in fused or sc > fused[nid]["score"]:          # duplicate ids keep the max (retrieval_backend.py:336-348)
                fused[nid] = {"id": nid, "score": sc, "meta": meta, "_row": r}
            if not use_text and len(fused) >= top_k:
                break
        if use_text:
            # fusion with the host text channel: retrieval_backend.py:336-372 on row-joined ids
            tmap: Dict[str, Dict[str, Any]] = {}
            for h in t_hits_raw:
                nid, meta = be._normalize_id(h)
                sc = float(h.get("score") or 0.0)
                if nid not in tmap or sc > tmap[nid]["score"]:
                    tmap[nid] = {"score": sc, "meta": dict(meta or {})}
            norm_text = be._minmax_norm({k: v["score"] for k, v in tmap.items()})
            for nid, ts in norm_text.items():
                if nid in fused:
                    fused[nid]["score"] += be.alpha_text * ts
                    fused[nid]["meta"]["score_text_norm"] = ts
                else:
                    meta = dict(tmap[nid]["meta"])
                    meta["score_text_norm"], meta["score_graph_norm"], meta["score_dense_norm"] = ts, 0.0, 0.0
                    fused[nid] = {"id": nid, "score": float(be.alpha_text * ts), "meta": meta, "_row": 1 << 62}
        if c.get("mmr") and not use_text:
            ordered = list(fused.values())                                     # MMR pick order (dicts keep insertion order)
        else:
            ordered = sorted(fused.values(), key=lambda h: h["_row"])          # enumerate in id order ...
            ordered.sort(key=lambda h: h["score"], reverse=True)               # ... so the stable sort breaks ties by lower id
        hits = [{"id": h["id"], "score": h["score"], "meta": h["meta"]} for h in ordered[:top_k]]
        diagnostics = self._diagnostics(i, top_k, len(t_hits_raw), c["ncand"], len(c["rows"]))
        return {"hits": hits, "diagnostics": diagnostics}

    def output(self, i: int) -> RetrievalOut:
        # the hit dicts go to pydantic as they are: its compiled validator builds the Hit objects in one call, which is
        # faster than a Python-level Hit(...) per hit (and than model_construct, a Python loop)
        res = self.result(i)
        return RetrievalOut(hits=res["hits"], diagnostics=res["diagnostics"])
