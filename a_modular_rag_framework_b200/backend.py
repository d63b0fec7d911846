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

What changes underneath: instead of embedding <= bm25_pool_k BM25 candidates per query through the
provider and scoring them with a pure-Python cosine, the query embedding is searched EXACTLY against the
whole resident passage-embedding corpus on the GPU (DenseIndex), the top `seed_k` passages seed the CSR
neighbour expansion, and candidates are re-scored and fused on the device (expand_rescore).  The optional
text channel stays on the host and is fused with the reference's formula.

The compute engine is injectable (tests drive the host logic with an oracle engine); the default engine is
the CUDA library and raises if it is unavailable -- there is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import logging
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

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


class CudaEngine:
    """Default engine: the corpus and the graph live on one GPU; every call is a libb200r kernel."""

    def __init__(self, emb_path: str, graph_path: str, device: str, metric: str):
        import numpy as np
        import torch
        from .index import DenseIndex, PassageGraph

        if not torch.cuda.is_available():
            raise RuntimeError("B200RetrievalBackend needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device(device)
        X, meta, scales = store.load_embeddings(emb_path)
        tdtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32, "i8": torch.int8}[meta["dtype"]]
        host = torch.from_numpy(np.array(X, copy=True))   # the store is a read-only memmap
        if meta["dtype"] == "bf16":
            host = host.view(torch.bfloat16)
        self.X = host.to(self.device)
        self.scales = torch.from_numpy(scales).to(self.device) if scales is not None else None
        self.index = DenseIndex(self.X, metric=metric or meta.get("metric", "cos"), row_scale=self.scales)
        self.dtype = tdtype
        self.N, self.D = self.index.N, self.index.D
        self.graph = None
        if graph_path and os.path.exists(graph_path):
            row_ptr, col_idx = store.load_csr(graph_path)
            self.graph = PassageGraph(torch.from_numpy(row_ptr).to(self.device), torch.from_numpy(col_idx).to(self.device))

    def _q(self, qvecs: Sequence[Sequence[float]]):
        q = self.torch.tensor(qvecs, dtype=self.torch.float32)
        return q.to(self.index.query_dtype).to(self.device).contiguous()

    def dense_search(self, qvecs, k: int) -> Tuple[List[List[float]], List[List[int]]]:
        s, i = self.index.search(self._q(qvecs), k)
        return s.cpu().tolist(), i.cpu().tolist()

    def has_graph(self) -> bool:
        return self.graph is not None

    def expand_rescore(self, qvecs, seeds: List[List[int]], window: int, alpha_dense: float, alpha_graph: float, k: int) -> Dict[str, list]:
        from .index import expand_rescore
        seeds_t = self.torch.tensor(seeds, dtype=self.torch.int64, device=self.device)
        out = expand_rescore(self.index, self.graph, self._q(qvecs), seeds_t, window, alpha_dense, alpha_graph, k)
        return {key: v.cpu().tolist() for key, v in out.items()}


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
    graph_path: str = ""                            # store.save_csr file of the corpus-level next_in_doc graph
    seed_k: int = 20                                # dense top-S passages that seed the graph expansion
    device: str = "cuda:0"
    metric: str = "cos"
    text_searcher: Any = None                       # optional object with .search(queries=, top_k=) -> raw hits (host BM25)
    expander: Any = None                            # optional object with .expand(query=, trace_id=) -> List[str]
    engine: Any = None                              # injectable compute engine (tests); default CudaEngine

    docs: List[Dict[str, Any]] = field(default_factory=list, repr=False)

    def __post_init__(self):
        if not self.docs:
            self.docs = store.load_docs(self.index_path)
        if self.engine is None:
            self.engine = CudaEngine(self.emb_path, self.graph_path, self.device, self.metric)
        n = getattr(self.engine, "N", None)
        if self.docs and n is not None and n != len(self.docs):
            raise ValueError(f"embedding rows ({n}) and docs.jsonl lines ({len(self.docs)}) disagree")
        logger.info("[B200RetrievalBackend] init: rows=%s seed_k=%d alpha_text=%s alpha_graph=%s alpha_dense=%s graph_window=%d",
                    n, self.seed_k, self.alpha_text, self.alpha_graph, self.alpha_dense, self.graph_window)

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

    def _embed_query(self, query: str, trace_id: str) -> Optional[List[float]]:
        """router.embed exactly as DenseReranker.score calls it (retrieval_backend.py:227-231): dict or list result."""
        cache = getattr(self, "_embed_cache", None)
        if cache is not None and query in cache:
            return cache[query]
        try:
            q_emb = self.router.embed(model_hint=self._resolve_embed_model(), texts=[query], require={"trace_id": trace_id})
            qv = (q_emb.get("vectors") if isinstance(q_emb, dict) else q_emb)[0]
            return [float(x) for x in qv]
        except Exception as e:  # noqa: BLE001
            logger.error("[B200RetrievalBackend] query embed error: %s", e, exc_info=True)
            return None

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

    # ------------------------------------------------------------------------------------------ the call
    def run(self, req: Any) -> Dict[str, Any]:
        trace_id = getattr(req, "trace_id", None) or "trace-demo"
        top_k = int(getattr(req, "top_k", None) or self.default_top_k)

        with _span("Backend/Expand", self.sink, trace_id):
            expanded = list(self.expander.expand(query=req.query, trace_id=trace_id)) if self.expander is not None else []
            queries = [req.query] + expanded

        with _span("Backend/TextSearch", self.sink, trace_id):
            t_hits_raw = list(self.text_searcher.search(queries=queries, top_k=max(top_k, self.bm25_pool_k))) if self.text_searcher is not None else []

        gw_override = getattr(req, "graph_window", None)
        window = max(0, int(gw_override if isinstance(gw_override, int) else self.graph_window))
        dense_rows: List[int] = []
        dense_scores: List[float] = []
        rows: List[int] = []
        fused_dev: List[float] = []
        gnorm: List[float] = []
        dnorm: List[float] = []
        hops: List[int] = []
        cosv: List[float] = []
        ncand = 0
        qv = None
        with _span("Backend/DenseRerank", self.sink, trace_id):
            qv = self._embed_query(req.query, trace_id)
            dim = getattr(self.engine, "D", None)
            if qv is not None and (dim is None or len(qv) == dim) and any(qv):
                pool = max(self.seed_k, top_k)
                s, i = self.engine.dense_search([qv], pool)
                keep = [(sc, r) for sc, r in zip(s[0], i[0]) if r >= 0]
                dense_scores = [sc for sc, _ in keep]
                dense_rows = [r for _, r in keep]
            elif qv is not None:
                logger.warning("[B200RetrievalBackend] query embedding unusable (dim=%s, corpus dim=%s): dense channel empty",
                               len(qv), dim)

        use_graph = bool(getattr(req, "graph_id", "") or "") and self.engine.has_graph() and bool(dense_rows) and self.alpha_graph != 0.0
        with _span("Backend/GraphExpand", self.sink, trace_id):
            if use_graph:
                seeds = dense_rows[: self.seed_k]
                out = self.engine.expand_rescore([qv], [seeds], window, self.alpha_dense, self.alpha_graph, max(top_k, self.bm25_pool_k))
                ncand = int(out["ncand"][0])
                for j, r in enumerate(out["id"][0]):
                    if r < 0:
                        break
                    rows.append(int(r)); fused_dev.append(float(out["score"][0][j])); gnorm.append(float(out["score_graph_norm"][0][j]))
                    dnorm.append(float(out["score_dense_norm"][0][j])); hops.append(int(out["hop"][0][j])); cosv.append(float(out["cos"][0][j]))
            else:
                # dense-only: the candidate set is the exact dense top-k; min-max over it like the reference does over its pool
                nd = self._minmax_norm({str(r): sc for r, sc in zip(dense_rows, dense_scores)})
                for r, sc in zip(dense_rows, dense_scores):
                    rows.append(r); dnorm.append(nd[str(r)]); gnorm.append(0.0); hops.append(-1); cosv.append(sc)
                    fused_dev.append(self.alpha_dense * nd[str(r)])

        # ---- fusion with the (optional) host text channel: retrieval_backend.py:336-372 on row-joined ids ----
        fused: Dict[str, Dict[str, Any]] = {}
        for j, r in enumerate(rows):
            meta = self._row_meta(r)
            meta["source"] = "graph" if (use_graph and hops[j] > 0) else "dense"
            if hops[j] >= 0:
                meta["distance"] = hops[j]
            meta["score_dense"] = cosv[j]
            nid, _ = self._normalize_id({"id": f"sent::row{r}", "meta": meta})
            meta["score_text_norm"], meta["score_graph_norm"], meta["score_dense_norm"] = 0.0, gnorm[j], dnorm[j]
            if nid not in fused or fused_dev[j] > fused[nid]["score"]:
                fused[nid] = {"id": nid, "score": float(fused_dev[j]), "meta": meta, "_row": r}
        if t_hits_raw and self.alpha_text != 0.0:
            tmap: Dict[str, Dict[str, Any]] = {}
            for h in t_hits_raw:
                nid, meta = self._normalize_id(h)
                sc = float(h.get("score") or 0.0)
                if nid not in tmap or sc > tmap[nid]["score"]:
                    tmap[nid] = {"score": sc, "meta": dict(meta or {})}
            norm_text = self._minmax_norm({k: v["score"] for k, v in tmap.items()})
            for nid, ts in norm_text.items():
                if nid in fused:
                    fused[nid]["score"] += self.alpha_text * ts
                    fused[nid]["meta"]["score_text_norm"] = ts
                else:
                    meta = dict(tmap[nid]["meta"])
                    meta["score_text_norm"], meta["score_graph_norm"], meta["score_dense_norm"] = ts, 0.0, 0.0
                    fused[nid] = {"id": nid, "score": float(self.alpha_text * ts), "meta": meta, "_row": 1 << 62}
        ordered = sorted(fused.values(), key=lambda h: h["_row"])          # enumerate in id order ...
        ordered.sort(key=lambda h: h["score"], reverse=True)               # ... so the stable sort breaks ties by lower id
        hits = [{"id": h["id"], "score": h["score"], "meta": h["meta"]} for h in ordered[:top_k]]

        diagnostics = {
            "queries": queries,
            "bm25_candidates": len(t_hits_raw),
            "graph_candidates": ncand,
            "dense_scored": getattr(self.engine, "N", len(dense_rows)) if dense_rows else 0,
            "weights": {"alpha_text": self.alpha_text, "alpha_graph": self.alpha_graph, "alpha_dense": self.alpha_dense},
            "pool": {"bm25_pool_k": self.bm25_pool_k, "final_top_k": top_k, "seed_k": self.seed_k},
            "graph_window_used": window,
            "embed_batch": int(self.embed_batch),
            "resolved_embed_model": self._resolve_embed_model(),
            "engine": type(self.engine).__name__,
        }
        return {"hits": hits, "diagnostics": diagnostics}

    def retrieve_many(self, reqs: Sequence[Any]) -> List[RetrievalOut]:
        """Batching shim: the orchestrator issues one query at a time
        (app/orchestrator/nodes.py:110-115, my_code/run_system.py:35-39); a batch runner can hand all
        requests over at once so the dense stage is ONE search launch and the graph stage ONE expand launch.
        Results are identical to calling retrieve() per request (same engine calls, batched)."""
        if not reqs:
            return []
        self._embed_cache = {}
        batch = _BatchedEngine(self, reqs)
        saved = self.engine
        self.engine = batch
        try:
            return [self.retrieve(r) for r in reqs]
        finally:
            self.engine = saved
            self._embed_cache = None

    def retrieve(self, req: Any) -> RetrievalOut:
        result = self.run(req)
        hits = [Hit(id=h["id"], score=h["score"], meta=h.get("meta")) for h in result["hits"]]
        return RetrievalOut(hits=hits, diagnostics=result.get("diagnostics", {}))


class _BatchedEngine:
    """Engine proxy used by retrieve_many: pre-computes the dense search (and, when every request takes the
    graph path with the same window, the expansion) for the whole batch, then serves the per-request calls."""

    def __init__(self, be: "B200RetrievalBackend", reqs: Sequence[Any]):
        self.inner = be.engine
        self.N = getattr(self.inner, "N", None)
        self.D = getattr(self.inner, "D", None)
        self._dense: Dict[Tuple[float, ...], Tuple[List[float], List[int]]] = {}
        self._graph: Dict[Tuple[Tuple[float, ...], Tuple[int, ...], int], Dict[str, list]] = {}
        qvs = []
        for r in reqs:
            qv = be._embed_query(r.query, getattr(r, "trace_id", None) or "trace-demo")
            be._embed_cache[r.query] = qv
            if qv is not None and (self.D is None or len(qv) == self.D) and any(qv):
                qvs.append(qv)
        if not qvs:
            return
        pools = {max(be.seed_k, int(getattr(r, "top_k", None) or be.default_top_k)) for r in reqs}
        pool = max(pools)
        s, i = self.inner.dense_search(qvs, pool)
        for qv, ss, ii in zip(qvs, s, i):
            self._dense[tuple(qv)] = (ss, ii)
        self._pool = pool
        windows = {max(0, int(getattr(r, "graph_window", None) if isinstance(getattr(r, "graph_window", None), int) else be.graph_window)) for r in reqs}
        all_graph = all(bool(getattr(r, "graph_id", "") or "") for r in reqs) and self.inner.has_graph() and be.alpha_graph != 0.0
        if all_graph and len(windows) == 1 and len(pools) == 1:
            w = windows.pop()
            kk = max(pool, be.bm25_pool_k)
            seeds = [[r for r in self._dense[tuple(qv)][1] if r >= 0][: be.seed_k] for qv in qvs]
            S = max(len(sd) for sd in seeds)
            padded = [sd + [-1] * (S - len(sd)) for sd in seeds]
            out = self.inner.expand_rescore(qvs, padded, w, be.alpha_dense, be.alpha_graph, kk)
            for j, (qv, sd) in enumerate(zip(qvs, seeds)):
                self._graph[(tuple(qv), tuple(sd), w)] = {key: [v[j]] for key, v in out.items()}

    def has_graph(self) -> bool:
        return self.inner.has_graph()

    def dense_search(self, qvecs, k: int):
        hit = self._dense.get(tuple(qvecs[0])) if len(qvecs) == 1 else None
        if hit is not None and k <= self._pool:
            return [hit[0][:k]], [hit[1][:k]]
        return self.inner.dense_search(qvecs, k)

    def expand_rescore(self, qvecs, seeds, window, alpha_dense, alpha_graph, k):
        hit = self._graph.get((tuple(qvecs[0]), tuple(seeds[0]), window)) if len(qvecs) == 1 else None
        if hit is not None and k <= len(hit["id"][0]):
            return {key: ([v[0][:k]] if isinstance(v[0], list) else v) for key, v in hit.items()}
        return self.inner.expand_rescore(qvecs, seeds, window, alpha_dense, alpha_graph, k)
