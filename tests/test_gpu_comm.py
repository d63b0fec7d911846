"""The C-ABI shard exchange (b200r_comm_*, csrc/comm.cu): protocol tests with several VIRTUAL ranks on one device (same
code path: peer stores into windows, epoch flags, rank-merge kernel), the single-process multi-GPU host call, and a real
one-process-per-GPU run over CUDA IPC when two GPUs are visible."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as O

pytestmark = pytest.mark.gpu


def _shards(X, G):
    from a_modular_rag_framework_b200 import DenseIndex
    from a_modular_rag_framework_b200.sharded import shard_rows
    out = []
    for r in range(G):
        lo, hi = shard_rows(X.shape[0], G, r)
        out.append(DenseIndex(X[lo:hi], id_base=lo))
    return out


@pytest.mark.parametrize("G,N,D,B,k,dtype", [
    (4, 50001, 64, 24, 100, torch.bfloat16),
    (3, 20000, 128, 300, 10, torch.bfloat16),     # CTA pairs, odd B*k (record stride is padded to 16 bytes)
    (8, 9000, 64, 5, 33, torch.float16),
    (2, 400, 32, 3, 300, torch.bfloat16),         # k > rows per shard: padded lists
    (3, 30000, 96, 70, 20, torch.float32),        # scan path, sliced batch: record staged then scattered
])
def test_virtual_ranks_exchange_equals_unsharded(G, N, D, B, k, dtype):
    from a_modular_rag_framework_b200 import DenseIndex
    from a_modular_rag_framework_b200.comm import ShardComm
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(G * 1000 + B)
    X = torch.randn((N, D), generator=g, device=dev).to(dtype)
    X[N - 1] = X[7]                     # exact tie across shards: lower GLOBAL id first
    Q = torch.randn((B, D), generator=g, device=dev).to(dtype)
    Q[0] = X[7]
    full_s, full_i = DenseIndex(X).search(Q, k)
    shards = _shards(X, G)
    comms = ShardComm.init_all([0] * G, max_B=B, max_k=k)
    for rep in range(5):                 # several epochs through both slots
        for r in range(G):
            shards[r].search_sharded_begin(comms[r], Q, k)
        outs = [comms[r].finish(B, k) for r in range(G)]
        for s, i in outs:
            assert torch.equal(i, full_i)
            assert torch.equal(s, full_s)
    for c in comms:
        c.check()
    if G == 4:
        # the whole sharded step (every rank's begin, then every rank's finish) captured into ONE CUDA graph: the exchange
        # keeps its epochs on the device, so replays stay in step
        graph = torch.cuda.CUDAGraph()
        Qs = Q.clone()
        with torch.cuda.graph(graph):
            for r in range(G):
                shards[r].search_sharded_begin(comms[r], Qs, k)
            outs_g = [comms[r].finish(B, k) for r in range(G)]
        for rep in range(4):
            Qs.copy_(Q if rep % 2 == 0 else Q.flip(0))
            graph.replay()
            want_i = full_i if rep % 2 == 0 else full_i.flip(0)
            for s, i in outs_g:
                assert torch.equal(i, want_i)
        for c in comms:
            c.check()
    pos7, posl = (full_i[0] == 7).nonzero().item(), (full_i[0] == N - 1).nonzero().item()
    assert posl == pos7 + 1
    S = O.score_matrix(Q[:2].float().cpu().numpy(), X.float().cpu().numpy())
    tol = 1e-5 if dtype == torch.float32 else 1e-3
    for b in range(2):
        O.check_topk(outs[0][0][b].cpu().numpy(), outs[0][1][b].cpu().numpy(), S[b], k, rel_tol=tol)


def test_pipelined_begin_before_finish_and_capacity_errors():
    """Throughput order: begin(i+1) is issued before finish(i) (two slots in flight); a batch beyond the window's capacity
    is refused with a message, never silently truncated."""
    from a_modular_rag_framework_b200 import B200RError, DenseIndex
    from a_modular_rag_framework_b200.comm import ShardComm
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(11)
    G, N, D, B, k = 3, 30000, 64, 40, 20
    X = torch.randn((N, D), generator=g, device=dev).to(torch.bfloat16)
    Qs = [torch.randn((B, D), generator=g, device=dev).to(torch.bfloat16) for _ in range(6)]
    ref = [DenseIndex(X).search(q, k) for q in Qs]
    shards = _shards(X, G)
    comms = ShardComm.init_all([0] * G, max_B=B, max_k=k)
    res = []
    for i, q in enumerate(Qs):
        for r in range(G):
            shards[r].search_sharded_begin(comms[r], q, k)
        if i > 0:
            res.append([comms[r].finish(B, k) for r in range(G)])
    res.append([comms[r].finish(B, k) for r in range(G)])
    for (rs, ri), outs in zip(ref, res):
        for s, i in outs:
            assert torch.equal(i, ri) and torch.equal(s, rs)
    with pytest.raises(B200RError):
        shards[0].search_sharded_begin(comms[0], torch.cat([Qs[0], Qs[1]]), k)      # B = 80 > max_B = 40
    for c in comms:
        c.check()


def test_single_process_host_all_call():
    """b200r_search_sharded_host_all: host queries in, host results out, G device shards driven by one thread."""
    from a_modular_rag_framework_b200 import DenseIndex, _native as N_
    from a_modular_rag_framework_b200.comm import ShardComm
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(5)
    G, N, D, B, k = 4, 40000, 64, 130, 10
    X = torch.randn((N, D), generator=g, device=dev).to(torch.bfloat16)
    Q = torch.randn((B, D), generator=g, device=dev).to(torch.bfloat16)
    rs, ri = DenseIndex(X).search(Q, k)
    shards = _shards(X, G)
    comms = ShardComm.init_all([0] * G, max_B=B, max_k=k)
    Qh = Q.cpu().pin_memory()
    hs = torch.empty((B, k), dtype=torch.float32).pin_memory()
    hi = torch.empty((B, k), dtype=torch.int64).pin_memory()
    lib = N_.load()
    ih = (ctypes.c_uint64 * G)(*[s.handle for s in shards])
    ch = (ctypes.c_uint64 * G)(*[c.handle for c in comms])
    for _ in range(3):
        N_.check(lib.b200r_search_sharded_host_all(ih, ch, G, Qh.data_ptr(), B, k, 0, hs.data_ptr(), hi.data_ptr(), None))
        assert torch.equal(hi, ri.cpu()) and torch.equal(hs, rs.cpu())
    # fp32 host queries (provider embeddings), rounded to bf16 on every device: same result as the host-rounded ones
    Q32 = torch.randn((B, D), generator=g, device=dev)
    rs, ri = DenseIndex(X).search(Q32.to(torch.bfloat16), k)
    Q32h = Q32.cpu().pin_memory()
    hs.fill_(0); hi.fill_(0)
    N_.check(lib.b200r_search_sharded_host_all_f32(ih, ch, G, Q32h.data_ptr(), B, k, 0, hs.data_ptr(), hi.data_ptr(), None))
    assert torch.equal(hi, ri.cpu()) and torch.equal(hs, rs.cpu())


def _ipc_worker(rank, world, port, tmpdir, ret):
    import torch.distributed as dist
    from a_modular_rag_framework_b200 import DenseIndex
    from a_modular_rag_framework_b200.comm import ShardComm
    from a_modular_rag_framework_b200.sharded import ShardedDenseIndex, shard_rows
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(77)
    N, D, B, k = 30000, 64, 10, 20
    Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((N, D)).astype(np.float32)))
    Qn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((B, D)).astype(np.float32)))
    lo, hi = shard_rows(N, world, rank)
    local = DenseIndex(torch.from_numpy(Xn[lo:hi]).to(dev).to(torch.bfloat16), id_base=lo)
    Q = torch.from_numpy(Qn).to(dev).to(torch.bfloat16)
    S = O.score_matrix(Qn, Xn)
    # (1) no torch.distributed at all: descriptors through a directory
    comm = ShardComm.from_files(tmpdir, rank, rank, world, max_B=B, max_k=k)
    for _ in range(4):
        local.search_sharded_begin(comm, Q, k)
        s, i = comm.finish(B, k)
    for b in range(B):
        O.check_topk(s[b].cpu().numpy(), i[b].cpu().numpy(), S[b], k, rel_tol=1e-3)
    comm.check()
    # (2) through ShardedDenseIndex with torch.distributed as the descriptor channel; pipelined batches
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sh = ShardedDenseIndex(local, exchange="abi")
        s2, i2 = sh.search(Q, k)
        assert torch.equal(i2, i) and torch.equal(s2, s)
        pend = [sh.search(Q, k, overlap=True) for _ in range(5)]
        for pr in pend:
            s3, i3 = pr.get()
            assert torch.equal(i3, i) and torch.equal(s3, s)
        s4, i4 = ShardedDenseIndex(local, exchange="nccl").search(Q, k)
        assert torch.equal(i4, i) and torch.equal(s4, s)
        sh.comm.check()
        # graph stage over the exchange, one rank per process: home / owner sharding, all phases in one call
        from a_modular_rag_framework_b200 import PassageGraph
        from a_modular_rag_framework_b200.index import expand_rescore_exchange
        from tests.helpers_graph import check_graph_query
        row_ptr, col_idx = O.csr_from_edges(N, rng.integers(0, N, 4 * N), rng.integers(0, N, 4 * N))
        graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
        gcomm = ShardComm.from_process_group(rank, B, k, max_graph_k=10)
        seeds = i[:, :8].contiguous()
        for _ in range(3):
            out = expand_rescore_exchange(local, graph, gcomm, Q, seeds, 2, 0.4, 0.2, 10, N)
        for b in range(0, B, 3):
            check_graph_query(out, b, Qn[b], Xn, row_ptr, col_idx, seeds[b].cpu().numpy(), 2, 0.4, 0.2, 10)
        gcomm.check()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_processes_over_cuda_ipc(tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, str(tmp_path), ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


@pytest.mark.parametrize("G,B,V,E,D,S,k,window", [
    (4, 16, 30000, 120000, 64, 20, 10, 2),
    (3, 50, 20000, 60000, 96, 8, 25, 2),
    (8, 5, 9000, 30000, 64, 6, 8, 3),          # fewer queries than ranks: some ranks have no home query
    (2, 33, 5000, 0, 32, 4, 6, 2),             # no edges at all: candidates == seeds
])
def test_graph_stage_over_the_exchange_equals_single_gpu(G, B, V, E, D, S, k, window):
    """b200r_expand_rescore_exchange: queries sharded by home rank, candidates by owner rank, cosines returned to the home
    rank, results stored into every rank's window -- must reproduce the single-launch stage and the oracle."""
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore
    from a_modular_rag_framework_b200.comm import ShardComm
    from a_modular_rag_framework_b200.index import expand_rescore_exchange_all
    from tests.helpers_graph import check_graph_query
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(G * 100 + B)
    row_ptr, col_idx = O.csr_from_edges(V, rng.integers(0, V, E), rng.integers(0, V, E))
    N = V - 300                                               # the last 300 nodes have no corpus row (score 0.0)
    Xn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((N, D)).astype(np.float32)))
    Qn = O.bf16_bits_to_f32(O.f32_to_bf16_bits(rng.standard_normal((B, D)).astype(np.float32)))
    seeds = rng.integers(0, V, size=(B, S))
    seeds[:, -1] = -1
    X = torch.from_numpy(Xn).to(dev).to(torch.bfloat16)
    Q = torch.from_numpy(Qn).to(dev).to(torch.bfloat16)
    graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
    ref = expand_rescore(DenseIndex(X), graph, Q, torch.from_numpy(seeds).to(dev), window, 0.4, 0.2, k)
    shards = _shards(X, G)
    comms = ShardComm.init_all([0] * G, max_B=B, max_k=k, max_graph_k=k)
    for rep in range(3):
        out = expand_rescore_exchange_all(shards, [graph] * G, comms, N, Q.cpu(), torch.from_numpy(seeds), window, 0.4, 0.2, k)
        for key in ("id", "hop", "ncand", "status"):
            assert torch.equal(out[key], ref[key]), key
        assert torch.equal(out["cos"], ref["cos"])
        assert torch.allclose(out["score"], ref["score"], rtol=0, atol=1e-6)
        assert torch.allclose(out["score_dense_norm"], ref["score_dense_norm"], rtol=0, atol=1e-6)
    for c in comms:
        c.check()
    Xfull = np.concatenate([Xn, np.zeros((300, D), np.float32)])
    for b in range(0, B, 7):
        check_graph_query(out, b, Qn[b], Xfull, row_ptr, col_idx, seeds[b], window, 0.4, 0.2, k)


def test_graph_exchange_overflow_and_mapped_graphs():
    """A query whose frontier overflows is reported per query (status) to every rank while the others still get results;
    mapped graphs (local node ids -> corpus rows) work through the exchange."""
    from a_modular_rag_framework_b200 import DenseIndex, PassageGraph, expand_rescore
    from a_modular_rag_framework_b200.comm import ShardComm
    from a_modular_rag_framework_b200.index import expand_rescore_exchange_all
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    V, D, G, k = 20000, 32, 3, 10
    star_s, star_d = np.zeros(V - 1, dtype=np.int64), np.arange(1, V, dtype=np.int64)          # node 0 reaches everything
    row_ptr, col_idx = O.csr_from_edges(V, star_s, star_d)
    X = torch.from_numpy(rng.standard_normal((V, D)).astype(np.float32)).to(dev).to(torch.bfloat16)
    Q = torch.from_numpy(rng.standard_normal((4, D)).astype(np.float32)).to(dev).to(torch.bfloat16)
    seeds = np.array([[5], [0], [7], [9]])                                                    # query 1 overflows at window 1
    graph = PassageGraph(torch.from_numpy(row_ptr).to(dev), torch.from_numpy(col_idx).to(dev))
    comms = ShardComm.init_all([0] * G, max_B=8, max_k=k, max_graph_k=k)
    shards = _shards(X, G)
    out = expand_rescore_exchange_all(shards, [graph] * G, comms, V, Q.cpu(), torch.from_numpy(seeds), 1, 0.4, 0.2, k)
    ref = expand_rescore(DenseIndex(X), graph, Q, torch.from_numpy(seeds).to(dev), 1, 0.4, 0.2, k, check_status=False)
    assert out["status"].tolist() == ref["status"].tolist() and out["status"][1] != 0 and out["status"][0] == 0
    assert torch.equal(out["id"], ref["id"]) and (out["id"][1] == -1).all()
    # mapped graph: 3 local nodes in a chain whose corpus rows are far apart (different owner ranks)
    rp, ci = O.csr_from_edges(3, np.array([0, 1]), np.array([1, 2]))
    node_rows = np.array([19000, 40, 9000], dtype=np.int64)
    mg = PassageGraph(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(node_rows).to(dev))
    sd = np.array([[0], [2], [1], [-1]])
    out = expand_rescore_exchange_all(shards, [mg] * G, comms, V, Q.cpu(), torch.from_numpy(sd), 1, 0.4, 0.2, k)
    ref = expand_rescore(DenseIndex(X), mg, Q, torch.from_numpy(sd).to(dev), 1, 0.4, 0.2, k)
    assert torch.equal(out["id"], ref["id"]) and torch.equal(out["hop"], ref["hop"]) and torch.equal(out["cos"], ref["cos"])
    assert sorted(out["id"][0][:2].tolist()) == [40, 19000] and out["ncand"].tolist() == [2, 2, 3, 0]
