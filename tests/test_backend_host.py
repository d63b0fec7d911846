"""Host logic of B200RetrievalBackend (CPU): boundary shape, ids/meta, fusion, spans, adapter round trip.
The compute engine is an oracle stand-in here; tests/test_gpu_backend.py runs the same flow on the CUDA engine."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from a_modular_rag_framework_b200 import store
from a_modular_rag_framework_b200.backend import B200RetrievalBackend
from tests.helpers_backend import FakeRouter, ListSink, OracleEngine, make_corpus, text_vec
from oracle import retrieval_oracle as O

REFERENCE = "/root/reference"


def _req(query, graph_id="g1", top_k=10, trace_id="t1", **kw):
    return SimpleNamespace(query=query, graph_id=graph_id, top_k=top_k, trace_id=trace_id, **kw)


@pytest.fixture()
def corpus(tmp_path):
    return make_corpus(str(tmp_path))


def _backend(c, **kw):
    eng = OracleEngine(c.X, c.row_ptr, c.col_idx)
    return B200RetrievalBackend(router=FakeRouter(c.dim), sink=kw.pop("sink", None), index_path=c.index_path,
                                emb_path=c.emb_path, graph_path=c.graph_path, engine=eng, **kw)


def test_store_roundtrip(corpus):
    X, meta, scales = store.load_embeddings(corpus.emb_path)
    assert meta == {"N": 60, "D": 64, "dtype": "bf16", "metric": "cos", "has_scales": False} and scales is None
    assert np.array_equal(np.asarray(X), O.f32_to_bf16_bits(corpus.X))
    assert store.load_docs(corpus.index_path) == corpus.docs
    rp, ci = store.load_csr(corpus.graph_path)
    assert np.array_equal(rp, corpus.row_ptr) and np.array_equal(ci, corpus.col_idx)
    # chain inside each 5-sentence document, nothing across documents
    assert rp[-1] == 2 * 12 * 4 and list(ci[rp[0]:rp[1]]) == [1] and list(ci[rp[2]:rp[3]]) == [1, 3]


def test_graph_json_to_csr_follows_build_index():
    graph = {"nodes": [], "edges": [
        {"source": "a", "target": "b", "type": "next_in_doc"}, {"source": "b", "target": "c", "type": "next_in_doc"},
        {"source": "a", "target": "c", "type": "semantic_sim"}, {"source": "q1", "target": "b", "type": "q_match"},
        {"source": "q2", "target": "a", "type": "q_match"}, {"source": "c", "target": "zz", "type": "next_in_doc"}]}
    rows = {"a": 0, "b": 1, "c": 2}
    rp, ci, seeds = store.graph_json_to_csr(graph, rows.get, 3)
    assert seeds == [1] and list(rp) == [0, 1, 3, 4] and list(ci) == [1, 0, 2, 1]


def test_run_shape_ids_meta_and_diagnostics(corpus):
    sink = ListSink()
    be = _backend(corpus, sink=sink, seed_k=5)
    q = corpus.docs[17]["text"]
    out = be.run(_req(q, top_k=8))
    hits, diag = out["hits"], out["diagnostics"]
    assert len(hits) == 8 and all(set(h) == {"id", "score", "meta"} for h in hits)
    assert [h["score"] for h in hits] == sorted((h["score"] for h in hits), reverse=True)
    top = hits[0]
    assert top["meta"]["row"] == 17 and top["meta"]["doc"] == "Title 3" and top["meta"]["sent_id"] == 2
    assert top["id"] == "sent::Title 3::2" and top["meta"]["text"] == q
    for key in ("score_text_norm", "score_graph_norm", "score_dense_norm", "kind", "text", "doc", "sent_id", "source"):
        assert key in top["meta"]
    # sentence 0 ids mirror the reference's falsy-sent_id behaviour (retrieval_backend.py:287-291)
    zero = [h for h in hits if h["meta"]["sent_id"] == 0]
    assert all(h["id"].endswith("::") for h in zero)
    for key in ("queries", "bm25_candidates", "graph_candidates", "dense_scored", "weights", "pool", "graph_window_used",
                "embed_batch", "resolved_embed_model"):
        assert key in diag
    assert diag["resolved_embed_model"] == "fake-embed" and diag["graph_window_used"] == 2 and diag["queries"] == [q]
    nodes = [e["node"] for e in sink.events if e["event"] == "node_end"]
    assert nodes == ["Backend/Expand", "Backend/TextSearch", "Backend/DenseRerank", "Backend/GraphExpand"]
    assert be.router.calls[0][0] == "fake-embed" and be.router.calls[0][2] == {"trace_id": "t1"}


def test_graph_stage_matches_oracle_composition(corpus):
    be = _backend(corpus, seed_k=4, graph_window=2)
    q = corpus.docs[31]["text"]
    out = be.run(_req(q, top_k=10))
    qv = O.bf16_bits_to_f32(O.f32_to_bf16_bits(text_vec(q, corpus.dim)))
    Xb = O.bf16_bits_to_f32(O.f32_to_bf16_bits(corpus.X))
    _, seeds = O.topk(qv[None], Xb, 10)
    top, dist = O.expand_rescore(qv, Xb, corpus.row_ptr, corpus.col_idx, seeds[0][:4], 2, 0.4, 0.2, 150)
    assert [h["meta"]["row"] for h in out["hits"]] == [t[0] for t in top[:10]]
    assert np.allclose([h["score"] for h in out["hits"]], [t[1] for t in top[:10]])
    assert out["diagnostics"]["graph_candidates"] == len(dist)
    assert all(h["meta"]["distance"] == dist[h["meta"]["row"]] for h in out["hits"])


def test_dense_only_when_no_graph_id_or_window_zero(corpus):
    be = _backend(corpus)
    q = corpus.docs[5]["text"]
    out = be.run(_req(q, graph_id="", top_k=5))
    qv = O.bf16_bits_to_f32(O.f32_to_bf16_bits(text_vec(q, corpus.dim)))
    s, i = O.topk(qv[None], O.bf16_bits_to_f32(O.f32_to_bf16_bits(corpus.X)), 20)
    assert [h["meta"]["row"] for h in out["hits"]] == list(i[0][:5])
    assert out["hits"][0]["meta"]["score_dense_norm"] == 1.0 and out["hits"][0]["score"] == pytest.approx(0.4)
    assert out["diagnostics"]["graph_candidates"] == 0
    out0 = be.run(_req(q, top_k=5, graph_window=0))      # per-request override, like getattr(req, "graph_window") (:318)
    assert out0["diagnostics"]["graph_window_used"] == 0 and all(h["meta"]["distance"] == 0 for h in out0["hits"])


def test_text_channel_fusion(corpus):
    class Text:
        def search(self, queries, top_k):
            return [{"id": "sent::Title 9::3", "score": 7.0, "meta": {"doc": "Title 9", "sent_id": 3, "text": "x", "source": "bm25"}},
                    {"id": "sent::Title 0::1", "score": 3.0, "meta": {"doc": "Title 0", "sent_id": 1, "text": "y", "source": "bm25"}}]
    be = _backend(corpus, text_searcher=Text())
    out = be.run(_req(corpus.docs[1]["text"], top_k=60))
    by_id = {h["id"]: h for h in out["hits"]}
    assert by_id["sent::Title 9::3"]["meta"]["score_text_norm"] == 1.0
    assert by_id["sent::Title 0::1"]["meta"]["score_text_norm"] == 0.0
    assert out["diagnostics"]["bm25_candidates"] == 2
    h = by_id["sent::Title 0::1"]   # joined on id with the dense/graph candidate for row 1
    assert h["meta"]["row"] == 1 and h["score"] == pytest.approx(0.2 * h["meta"]["score_graph_norm"] + 0.4 * h["meta"]["score_dense_norm"])


def test_degenerate_embeddings_degrade_like_reference(corpus):
    be = _backend(corpus)
    be.router = SimpleNamespace(policy={}, providers={}, embed=lambda **kw: [[0.0, 0.0, 0.0]])   # the router's mock fallback (llm_router.py:116-129)
    out = be.run(_req("anything"))
    assert out["hits"] == [] and out["diagnostics"]["dense_scored"] == 0

    def boom(**kw):
        raise RuntimeError("provider down")
    be.router = SimpleNamespace(policy={}, providers={}, embed=boom)
    assert be.run(_req("anything"))["hits"] == []


def test_retrieve_returns_retrievalout(corpus):
    be = _backend(corpus)
    ro = be.retrieve(_req(corpus.docs[3]["text"], top_k=3))
    assert len(ro.hits) == 3 and ro.hits[0].meta["row"] == 3 and "weights" in ro.diagnostics


def test_rows_docs_mismatch_raises(corpus):
    with pytest.raises(ValueError):
        B200RetrievalBackend(router=FakeRouter(corpus.dim), index_path=corpus.index_path, engine=OracleEngine(corpus.X[:10]))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "app")), reason="reference checkout not present on this box")
def test_through_the_references_own_adapter_and_flow_conventions(corpus, monkeypatch):
    """Plug-in convention B (retrieval_adapter.py:31-42, 112-133) drives our class by its import string, and
    convention A's kwarg filtering (flow.py:97-106) finds `router`/`sink` in our signature."""
    import inspect
    monkeypatch.syspath_prepend(REFERENCE)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    from app.modules.retrieval.retrieval_adapter import RetrievalAdapter
    from app.core.dto import RetrievalIn, RetrievalOut
    eng = OracleEngine(corpus.X, corpus.row_ptr, corpus.col_idx)
    ad = RetrievalAdapter(router=FakeRouter(corpus.dim), sink=None,
                          backend_impl="a_modular_rag_framework_b200.backend:B200RetrievalBackend",
                          backend_kwargs={"index_path": corpus.index_path, "emb_path": corpus.emb_path,
                                          "graph_path": corpus.graph_path, "engine": eng, "seed_k": 5})
    out = ad.retrieve(RetrievalIn(query=corpus.docs[42]["text"], graph_id="g", top_k=6, trace_id="t"))
    assert isinstance(out, RetrievalOut) and len(out.hits) == 6 and out.hits[0].meta["row"] == 42
    assert [h.score for h in out.hits] == sorted((h.score for h in out.hits), reverse=True)
    params = inspect.signature(B200RetrievalBackend.__init__).parameters
    for name in ("router", "sink", "index_path", "graph_root", "graph_window", "alpha_text", "alpha_graph", "alpha_dense",
                 "bm25_pool_k", "default_top_k", "qe_lines", "qe_attr_paraphrase", "embed_batch"):
        assert name in params


def test_retrieve_many_equals_retrieve(corpus):
    be = _backend(corpus, seed_k=4)
    reqs = [_req(corpus.docs[r]["text"], top_k=7, trace_id=f"t{r}") for r in (3, 14, 15, 42)]
    single = [be.retrieve(r) for r in reqs]
    many = be.retrieve_many(reqs)
    assert len(many) == 4
    for a, b in zip(single, many):
        assert [h.id for h in a.hits] == [h.id for h in b.hits]
        assert np.allclose([h.score for h in a.hits], [h.score for h in b.hits])
        assert a.diagnostics["graph_candidates"] == b.diagnostics["graph_candidates"]
    assert type(be.engine).__name__ == "OracleEngine"   # proxy removed afterwards
    mixed = be.retrieve_many([_req(corpus.docs[5]["text"], graph_id=""), _req(corpus.docs[6]["text"], top_k=3)])
    assert mixed[0].hits[0].meta["row"] == 5 and len(mixed[1].hits) == 3


def test_batch_converted_materialisation_equals_the_general_path(corpus, monkeypatch):
    """RetrievalBatch.result has two routes to the same hits: lists converted once per batch (no text channel, no MMR)
    and per-request arrays.  Graph and dense-only requests, top_k beyond the corpus, id collisions on sent_id 0
    (every document's first sentence normalises to 'sent::<doc>::', retrieval_backend.py:283-294)."""
    from a_modular_rag_framework_b200.backend import RetrievalBatch
    be = _backend(corpus, seed_k=6)
    reqs = [_req(corpus.docs[r]["text"], top_k=k, graph_id=g, trace_id=f"t{r}")
            for r, k, g in ((0, 10, "g1"), (5, 3, "g1"), (10, 10, ""), (17, 100, ""), (20, 100, "g1"), (31, 1, "g1"), (45, 25, ""))]
    batch = be.run_batch(reqs)
    fast = [batch.result(i) for i in range(len(reqs))]
    assert batch._heads, "the batch-converted route did not run"
    monkeypatch.setattr(RetrievalBatch, "_fast_hits", lambda self, i, top_k: None)
    slow_batch = be.run_batch(reqs)
    slow = [slow_batch.result(i) for i in range(len(reqs))]
    assert not slow_batch._heads
    for a, b in zip(fast, slow):
        assert a == b
    assert any(h["id"].endswith("::") for h in fast[0]["hits"])     # a sent_id-0 collision id is in play
    assert len(fast[3]["hits"]) <= len(corpus.docs)


def test_query_matrix_accepts_lists_and_arrays():
    qm = B200RetrievalBackend._query_matrix
    v = [[1.0, 2.0, 3.0], None, [0.0, 0.0, 0.0], [4.0, 5.0], [6.0, 7.0, 8.0]]
    valid, Q = qm(v, 3)
    assert valid == [0, 4] and Q.dtype == np.float32 and np.array_equal(Q, np.asarray([[1, 2, 3], [6, 7, 8]], dtype=np.float32))
    valid, Q = qm([np.asarray([1.0, 0.0]), np.asarray([0.0, 2.0])], None)
    assert valid == [0, 1] and Q.shape == (2, 2)
    assert qm([None, []], 3) == ([], None)


# ------------------------------------------------------------------------------------------------ a5: graph_id -> graph.json


def _bfs_cases(golden_dir):
    import json
    with open(os.path.join(golden_dir, "bfs_cases.json")) as f:
        return json.load(f)


def check_bfs_case_through_backend(be, case, gid):
    """The recorded output of the reference's expand_qmatch_neighbors (tests/golden/bfs_cases.json) must come back through
    the plug-in call: candidate set, hop distances (meta.distance) and the candidate count in the diagnostics."""
    out = be.run(_req("the question", graph_id=gid, top_k=max(1, len(case["nodes"])), graph_window=case["window"]))
    rows = sorted(h["meta"]["row"] for h in out["hits"])
    assert rows == case["nodes"]
    dist = dict(zip(case["nodes"], case["dist"]))
    assert all(h["meta"]["distance"] == dist[h["meta"]["row"]] for h in out["hits"])
    assert out["diagnostics"]["graph_candidates"] == len(case["nodes"])
    assert out["diagnostics"]["graph_window_used"] == case["window"]
    return out


def test_graph_id_graphs_replay_reference_bfs_cases(tmp_path, golden_dir):
    from tests.helpers_backend import make_graph_id_fixture
    for n, case in enumerate(_bfs_cases(golden_dir)):
        fx = make_graph_id_fixture(str(tmp_path / f"c{n}"), case, "g")
        eng = OracleEngine(fx.X)
        be = B200RetrievalBackend(router=FakeRouter(fx.dim), index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root,
                                  engine=eng)
        check_bfs_case_through_backend(be, case, "g")
        check_bfs_case_through_backend(be, case, "g")
        assert len(be._graphs) == 1                       # parsed once, not per query (retrieval_backend.py:154-155 does)
        # a graph_id without a graph.json and no corpus-level CSR: dense channel only, like the reference's empty graph
        out = be.run(_req("the question", graph_id="missing", top_k=5))
        assert out["diagnostics"]["graph_candidates"] == 0 and len(out["hits"]) == min(5, case["V"])


def test_graph_id_seed_fallbacks(tmp_path, golden_dir):
    """No q_match edges: the lexical-overlap seeds of graph_utils.py:76-83; no overlap either: the dense top-seed_k passages."""
    from tests.helpers_backend import make_graph_id_fixture
    case = _bfs_cases(golden_dir)[7]                      # V=60, window 2
    fx = make_graph_id_fixture(str(tmp_path), case, "g", qmatch=False)
    be = B200RetrievalBackend(router=FakeRouter(fx.dim), index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root,
                              engine=OracleEngine(fx.X), seed_k=3)
    lg = be.graph_for("g")
    assert lg.qmatch == [] and lg.V == 60
    # "alpha17" occurs in exactly one sentence -> that node is the only seed
    assert be._graph_seeds(lg, "tell me about alpha17", []) == [lg.row_to_node[17]]
    rp, ci = store.csr_from_edges(60, np.array(case["src"]), np.array(case["dst"]))
    want = O.bfs_expand(rp, ci, [17], 2)
    out = be.run(_req("tell me about alpha17", graph_id="g", top_k=60))
    assert sorted(h["meta"]["row"] for h in out["hits"]) == sorted(want)
    assert all(h["meta"]["distance"] == want[h["meta"]["row"]] for h in out["hits"])
    # no token in common with any sentence: seeds = dense top-3 rows
    out = be.run(_req("zzz qqq", graph_id="g", top_k=60))
    qv = O.bf16_bits_to_f32(O.f32_to_bf16_bits(text_vec("zzz qqq", fx.dim)))
    _, top = O.topk(qv[None], O.bf16_bits_to_f32(O.f32_to_bf16_bits(fx.X)), 3)
    want = O.bfs_expand(rp, ci, top[0].tolist(), 2)
    assert sorted(h["meta"]["row"] for h in out["hits"]) == sorted(want)


def test_batch_with_different_graph_ids_is_one_expansion(tmp_path, golden_dir):
    """retrieve_many over requests that name DIFFERENT graph_ids: one dense search and one expansion (block-diagonal union
    of the graphs), results equal to the one-request-at-a-time path."""
    from tests.helpers_backend import make_graph_id_fixture
    cases = _bfs_cases(golden_dir)
    case_a, case_b = cases[6], cases[7]                   # same V = 60 corpus, windows 1 and 2 -> same docs, two graphs
    fx = make_graph_id_fixture(str(tmp_path), case_a, "ga")
    # second graph over the same corpus: reversed chain edges and other seeds
    import json
    g2 = {"nodes": [], "edges": [{"source": f"Doc::sent{v}", "target": f"Doc::sent{v + 1}", "type": "next_in_doc"} for v in range(0, 59, 2)]
          + [{"source": "q1", "target": "Doc::sent10", "type": "q_match"}, {"source": "q1", "target": "Doc::sent41", "type": "q_match"}]}
    os.makedirs(os.path.join(fx.graph_root, "gb"))
    with open(os.path.join(fx.graph_root, "gb", "graph.json"), "w") as f:
        json.dump(g2, f)
    eng = OracleEngine(fx.X)
    be = B200RetrievalBackend(router=FakeRouter(fx.dim), index_path=fx.index_path, emb_path=fx.emb_path, graph_root=fx.graph_root, engine=eng)
    reqs = [_req("the question", graph_id="ga", top_k=20), _req("alpha3 sentence", graph_id="gb", top_k=20),
            _req("the question", graph_id="ga", top_k=7), _req("other", graph_id="", top_k=5), _req("x", graph_id="nope", top_k=5)]
    single = [be.retrieve(r) for r in reqs]
    eng.calls = {"dense": 0, "expand": 0}
    many = be.retrieve_many(reqs)
    assert eng.calls == {"dense": 1, "expand": 1}
    for a, b in zip(single, many):
        assert [h.id for h in a.hits] == [h.id for h in b.hits]
        assert np.allclose([h.score for h in a.hits], [h.score for h in b.hits])
    assert {h.meta["row"] for h in many[1].hits} == {10, 11, 40, 41}     # gb's edges are disjoint pairs (2v, 2v+1)


def test_mmr_post_filter(corpus):
    """mmr_lambda: the final top_k come from the best mmr_pool fused candidates in MMR pick order (similarity.py:33-62)."""
    plain = _backend(corpus, seed_k=5)
    be = _backend(corpus, seed_k=5, mmr_lambda=0.5, mmr_pool=30)
    q = corpus.docs[17]["text"]
    a, b = plain.run(_req(q, top_k=8)), be.run(_req(q, top_k=8))
    assert len(b["hits"]) == 8 and b["hits"][0]["meta"]["row"] == a["hits"][0]["meta"]["row"]      # first pick = best score
    pool = plain.run(_req(q, top_k=30))["hits"]
    Xb = O.bf16_bits_to_f32(O.f32_to_bf16_bits(corpus.X))
    items = [(h["meta"]["row"], np.float32(h["score"]).item(), Xb[h["meta"]["row"]].astype(float).tolist()) for h in pool]
    want = [p[0] for p in O.mmr_diversify(items, top_k=8, lambda_weight=0.5)]
    assert [h["meta"]["row"] for h in b["hits"]] == want
    many = be.retrieve_many([_req(q, top_k=8), _req(corpus.docs[3]["text"], top_k=4)])
    assert [h.meta["row"] for h in many[0].hits] == want and len(many[1].hits) == 4
