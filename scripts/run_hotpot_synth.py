"""BASELINE.json configs[0] stand-in: HotpotQA dev-distractor SHAPE (7,405 questions x 10 paragraphs, 384-d
embeddings, exact top-10) with synthetic text and seeded embeddings -- the real blobs are absent
(.MISSING_LARGE_BLOBS in the reference).  Runs every question through the reference-facing boundary in two ways and reports questions/s for both:
B200RetrievalBackend.retrieve_many (RetrievalOut / Hit objects for every request, what the orchestrator consumes) and
run_batch (the same computation, results left as arrays -- what a batch runner that only needs rows and scores reads).
A sample of questions is checked against the oracle engine.

    python scripts/run_hotpot_synth.py [--questions 7405] [--check 16]
"""
import argparse, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a_modular_rag_framework_b200 import store
from a_modular_rag_framework_b200.backend import B200RetrievalBackend
from tests.helpers_backend import FakeRouter, OracleEngine, text_vec
from types import SimpleNamespace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--questions", type=int, default=7405)
    ap.add_argument("--paragraphs", type=int, default=10)
    ap.add_argument("--sentences", type=int, default=4)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--check", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1024)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="hotpot_synth_")
    docs = []
    for q in range(a.questions):
        for p in range(a.paragraphs):
            title = f"Q{q} paragraph {p}"
            for s in range(a.sentences):
                docs.append({"doc_id": f"{title}#{s}", "title": title, "sent_id": s, "text": f"q{q} p{p} s{s}"})
    N = len(docs)
    index_path = os.path.join(tmp, "docs.jsonl")
    with open(index_path, "w") as f:
        for d in docs:
            f.write(json.dumps(d) + "\n")
    X = rng.standard_normal((N, a.dim)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    # make the gold sentence of question q (paragraph 0, sentence 0) resemble the question embedding
    qvecs = np.stack([text_vec(f"question {q}", a.dim) for q in range(a.questions)])
    gold = np.arange(a.questions) * a.paragraphs * a.sentences
    X[gold] = 0.7 * qvecs + 0.3 * X[gold]
    X[gold] /= np.linalg.norm(X[gold], axis=1, keepdims=True)
    from oracle import retrieval_oracle as O
    emb_path = os.path.join(tmp, "embeddings")
    store.save_embeddings(emb_path, O.f32_to_bf16_bits(X), "bf16")
    src, dst = store.next_in_doc_edges(docs)
    row_ptr, col_idx = store.csr_from_edges(N, src, dst)
    graph_path = os.path.join(tmp, "graph.npz")
    store.save_csr(graph_path, row_ptr, col_idx)

    class TableRouter(FakeRouter):
        """Embedding provider stand-in whose cost is a table lookup (the hash-seeded FakeRouter costs 20 us per text)."""
        def embed(self, model_hint, texts, require=None):
            return {"vectors": [table[t] for t in texts]}
    table = {f"question {q}": qvecs[q].tolist() for q in range(a.questions)}
    be = B200RetrievalBackend(router=TableRouter(a.dim), index_path=index_path, emb_path=emb_path, graph_path=graph_path, seed_k=20)
    reqs = [SimpleNamespace(query=f"question {q}", graph_id="g", top_k=10, trace_id=f"t{q}") for q in range(a.questions)]
    be.retrieve_many(reqs[:8])  # warm-up
    t0 = time.perf_counter()
    outs = []
    for i in range(0, len(reqs), a.batch):
        outs.extend(be.retrieve_many(reqs[i:i + a.batch]))
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    top1 = []
    for i in range(0, len(reqs), a.batch):
        batch = be.run_batch(reqs[i:i + a.batch])
        top1.extend(int(batch.rows(j)[0]) for j in range(len(batch.reqs)))
    dt_cols = time.perf_counter() - t0
    assert top1 == [o.hits[0].meta["row"] for o in outs]
    hit1 = float(np.mean([o.hits[0].meta["row"] == gold[q] for q, o in enumerate(outs)]))
    ref = B200RetrievalBackend(router=FakeRouter(a.dim), index_path=index_path, emb_path=emb_path, graph_path=graph_path, seed_k=20,
                               docs=be.docs, engine=OracleEngine(X, row_ptr, col_idx))
    agree = 0
    sample = rng.choice(a.questions, size=min(a.check, a.questions), replace=False)
    for q in sample:
        r = ref.retrieve(reqs[q])
        agree += [h.meta["row"] for h in r.hits] == [h.meta["row"] for h in outs[q].hits]
    print(json.dumps({"config": "hotpotqa-shaped synthetic (configs[0] stand-in)", "questions": a.questions, "rows": N, "dim": a.dim,
                      "top_k": 10, "questions_per_s_retrieve_many_hit_objects": a.questions / dt, "seconds": dt,
                      "questions_per_s_run_batch_arrays": a.questions / dt_cols, "seconds_arrays": dt_cols, "gold_at_1": hit1,
                      "oracle_agreement_on_sample": f"{agree}/{len(sample)}", "batch": a.batch}))


if __name__ == "__main__":
    main()
