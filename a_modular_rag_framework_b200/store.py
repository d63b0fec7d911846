"""On-disk formats either side of the hot path.

* ``docs.jsonl``  -- the reference's own sentence table (my_code/ingest_hotpotqa.py:73-81): one JSON
  object per line with doc_id/title/sent_id/text; line number == corpus row id
  (BM25LiteIndex._build, app/modules/retrieval/text_index.py:35-46).
* ``<dir>/embeddings.bin`` + ``<dir>/meta.json`` -- NEW: the row-aligned passage-embedding matrix
  (raw row-major, dtype bf16|f16|f32|i8, optional ``scales.f32.bin`` for i8).
* ``graph.json`` -> CSR -- the reference reloads and re-indexes graph.json per query
  (retrieval_backend.py:154-155, graph_utils.py:13-55); here it is converted once:
  next_in_doc edges only, both directions (graph_utils.py:49-51, 123), node id -> row id.
* ``<file>.npz`` -- CSR arrays row_ptr (int64) / col_idx (int32).
"""
from __future__ import annotations

import json
import os
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np

_NP_DTYPES = {"bf16": np.uint16, "f16": np.float16, "f32": np.float32, "i8": np.int8}


def load_docs(path: str) -> List[Dict[str, Any]]:
    """Same traversal as BM25LiteIndex._build: skip blank lines, row id = index in the returned list."""
    docs: List[Dict[str, Any]] = []
    if not os.path.exists(path):
        return docs
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            docs.append(json.loads(line))
    return docs


def save_embeddings(dirpath: str, X: np.ndarray, dtype: str, metric: str = "cos", scales: Optional[np.ndarray] = None) -> None:
    """X: [N, D] array already in the storage representation (uint16 bit patterns for bf16)."""
    if dtype not in _NP_DTYPES:
        raise ValueError(f"unknown dtype {dtype!r}")
    X = np.ascontiguousarray(X)
    if X.dtype != _NP_DTYPES[dtype] or X.ndim != 2:
        raise ValueError(f"X must be a 2-D {_NP_DTYPES[dtype].__name__} array for dtype {dtype}")
    os.makedirs(dirpath, exist_ok=True)
    X.tofile(os.path.join(dirpath, "embeddings.bin"))
    meta = {"N": int(X.shape[0]), "D": int(X.shape[1]), "dtype": dtype, "metric": metric, "has_scales": scales is not None}
    if scales is not None:
        np.ascontiguousarray(scales, dtype=np.float32).tofile(os.path.join(dirpath, "scales.f32.bin"))
    with open(os.path.join(dirpath, "meta.json"), "w", encoding="utf-8") as f:
        json.dump(meta, f)


def load_embeddings(dirpath: str) -> Tuple[np.ndarray, Dict[str, Any], Optional[np.ndarray]]:
    """-> (memory-mapped [N, D] array in storage representation, meta, scales or None)."""
    with open(os.path.join(dirpath, "meta.json"), "r", encoding="utf-8") as f:
        meta = json.load(f)
    X = np.memmap(os.path.join(dirpath, "embeddings.bin"), dtype=_NP_DTYPES[meta["dtype"]], mode="r",
                  shape=(meta["N"], meta["D"]))
    scales = None
    if meta.get("has_scales"):
        scales = np.fromfile(os.path.join(dirpath, "scales.f32.bin"), dtype=np.float32)
    return X, meta, scales


def graph_json_to_csr(graph: Dict[str, Any], node_to_row: Callable[[str], Optional[int]], num_rows: int) -> Tuple[np.ndarray, np.ndarray, List[int]]:
    """graph.json dict -> (row_ptr int64 [num_rows+1], col_idx int32, q_match seed rows).

    Follows build_index (graph_utils.py:46-53): only ``next_in_doc`` edges feed the adjacency (forward
    and backward), ``q_match`` edges whose source is "q1" give the seeds; every other edge type is ignored.
    Nodes that do not map to a corpus row are dropped."""
    src: List[int] = []
    dst: List[int] = []
    seeds: List[int] = []
    for e in graph.get("edges", []):
        et = e.get("type")
        s, t = e.get("source"), e.get("target")
        if et == "next_in_doc":
            a, b = node_to_row(s), node_to_row(t)
            if a is not None and b is not None:
                src.append(a)
                dst.append(b)
        elif et == "q_match" and s == "q1":
            r = node_to_row(t)
            if r is not None:
                seeds.append(r)
    row_ptr, col_idx = csr_from_edges(num_rows, np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64))
    return row_ptr, col_idx, seeds


def csr_from_edges(num_nodes: int, src: np.ndarray, dst: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Symmetrised CSR (neighbours = forward + backward), neighbour lists sorted by id."""
    a = np.concatenate([src, dst]).astype(np.int64)
    b = np.concatenate([dst, src]).astype(np.int64)
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    counts = np.bincount(a, minlength=num_nodes) if len(a) else np.zeros(num_nodes, dtype=np.int64)
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr, b.astype(np.int32)


def next_in_doc_edges(docs: Iterable[Dict[str, Any]]) -> Tuple[np.ndarray, np.ndarray]:
    """Corpus-level next_in_doc chain: consecutive sentences of the same title, as EdgeBuilder emits them
    per question (app/modules/graph_construction/edge_builder.py, `next_in_doc`)."""
    src, dst = [], []
    prev_title, prev_row = None, None
    for row, d in enumerate(docs):
        title = d.get("title")
        if prev_title is not None and title == prev_title:
            src.append(prev_row)
            dst.append(row)
        prev_title, prev_row = title, row
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)


def csr_path(path: str) -> str:
    """The file save_csr writes for `path`: numpy appends '.npz' to a name that lacks it."""
    return path if path.endswith(".npz") else path + ".npz"


def save_csr(path: str, row_ptr: np.ndarray, col_idx: np.ndarray) -> str:
    out = csr_path(path)
    np.savez(out, row_ptr=np.asarray(row_ptr, dtype=np.int64), col_idx=np.asarray(col_idx, dtype=np.int32))
    return out


def load_csr(path: str) -> Tuple[np.ndarray, np.ndarray]:
    z = np.load(path if os.path.exists(path) else csr_path(path))
    return z["row_ptr"].astype(np.int64), z["col_idx"].astype(np.int32)


def graph_json_to_local(graph: Dict[str, Any], node_to_row: Callable[[str], Optional[int]]
                        ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[int], List[str]]:
    """graph.json dict (one graph_id) -> LOCAL CSR + node -> corpus row map + q_match seeds + sentence texts.

    Follows build_index (graph_utils.py:25-55): ``next_in_doc`` edges feed the undirected adjacency, ``q_match`` edges
    whose source is "q1" give the seeds, sentence nodes contribute ``props.text`` (the lexical seed fallback of
    expand_qmatch_neighbors, :76-83); every other edge / node type is ignored.  Nodes that do not map to a corpus row
    cannot be scored or materialised and are dropped.  Local node ids follow first appearance (nodes, then edges).
    -> (row_ptr int64 [V+1], col_idx int32, node_rows int64 [V], qmatch local ids, texts [V])"""
    local: Dict[str, int] = {}
    rows: List[int] = []
    texts: List[str] = []

    def node(nid: Any) -> Optional[int]:
        if nid in local:
            return local[nid]
        r = node_to_row(nid) if isinstance(nid, str) else None
        if r is None:
            return None
        local[nid] = len(rows)
        rows.append(int(r))
        texts.append("")
        return local[nid]

    for n in graph.get("nodes", []):
        if n.get("type") == "sentence":
            v = node(n.get("id"))
            if v is not None:
                texts[v] = ((n.get("props") or {}).get("text") or "")
    src: List[int] = []
    dst: List[int] = []
    qmatch: List[int] = []
    for e in graph.get("edges", []):
        et = e.get("type")
        s, t = e.get("source"), e.get("target")
        if et == "next_in_doc":
            a, b = node(s), node(t)
            if a is not None and b is not None:
                src.append(a)
                dst.append(b)
        elif et == "q_match" and s == "q1":
            v = node(t)
            if v is not None and v not in qmatch:
                qmatch.append(v)
    V = len(rows)
    row_ptr, col_idx = csr_from_edges(V, np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)) if V else (
        np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32))
    return row_ptr, col_idx, np.asarray(rows, dtype=np.int64), qmatch, texts
