"""MCClassS forward (models/MCClassS.py topology) through the builder API on the GPU:
  * exact mode vs the same network assembled from the CPU oracle ops (indices bit-exact, logits to tolerance);
  * planned (sync-free) and CUDA-graph replays reproduce the exact-mode logits BIT FOR BIT (all kernels deterministic);
  * capacity overflow is detected and repaired, never silently truncated.
"""
import numpy as np
import pytest

from common import make_batch

pytestmark = pytest.mark.gpu


def _randomize_store(store, seed=0):
    """non-trivial biases / batch-norm statistics so that every parameter path is exercised"""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    for name, v in store.vars.items():
        if name.endswith("moving_variance"):
            v.copy_((torch.rand(v.shape, generator=g) * 0.5 + 0.75).to(v.device))
        elif "biases" in name or name.endswith("beta") or name.endswith("moving_mean"):
            v.copy_(((torch.rand(v.shape, generator=g) - 0.5) * 0.2).to(v.device))
        elif name.endswith("gamma"):
            v.copy_((torch.rand(v.shape, generator=g) * 0.5 + 0.75).to(v.device))


def _make(B, k, numCat, dev):
    import torch
    from mccnn_b200 import variables
    from mccnn_b200.models import MCClassS
    store = variables.VariableStore(device=dev, seed=3)

    def fn(p, b, f):
        return MCClassS.create_network(p, b, f, 1, B, k, numCat, False, 1.0, 1.0, store=store)
    return fn, store


@pytest.mark.parametrize("B,N,k", [(4, 1024, 16), (32, 1024, 64)])
def test_mcclasss_exact_planned_graph_identical(B, N, k):
    import torch
    from mccnn_b200.engine import PlannedForward
    dev = torch.device("cuda:0")
    fn, store = _make(B, k, 40, dev)
    pts, bids, feats = make_batch(B, N, "sphere", seed=11)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    ref = fn(tp, tb, tf)                      # creates the variables
    _randomize_store(store)
    ref = fn(tp, tb, tf)[:B].clone()          # exact mode
    assert ref.shape == (B, 40) and torch.isfinite(ref).all()

    # every kernel of the forward (Monte-Carlo ops and dense glue) is deterministic and independent of the capacity
    # padding, so planned / graph-replayed logits equal the exact-mode ones bit for bit
    def close(a, b):
        return torch.equal(a, b)

    eng_plain = PlannedForward(fn, B, use_graph=False)
    eng_graph = PlannedForward(fn, B, use_graph=True)
    for eng in (eng_plain, eng_graph):
        first = eng(tp, tb, tf).clone()       # records
        assert torch.equal(first, ref)
        outs = [eng(tp, tb, tf).clone() for _ in range(3)]  # planned / replayed
        assert close(outs[0], ref)
        assert torch.equal(outs[1], outs[0]) and torch.equal(outs[2], outs[0])  # replays are bit-reproducible
    assert torch.equal(eng_plain(tp, tb, tf), eng_graph(tp, tb, tf))           # graph == eager planned
    assert eng_graph.stats["captures"] == 1 and eng_graph.stats["replays"] == 4

    # a different batch of the same shape through the same captured graph
    pts2, bids2, feats2 = make_batch(B, N, "sphere", seed=12)
    tp2, tb2, tf2 = (torch.from_numpy(a).to(dev) for a in (pts2, bids2, feats2))
    ref2 = fn(tp2, tb2, tf2)[:B].clone()
    out2 = eng_graph(tp2, tb2, tf2).clone()
    assert close(out2, ref2)


def test_fused_neighbors_pdf_builder_path_is_bit_identical(monkeypatch):
    """MCCNN_FUSE_PDF=1 (one launch for find_neighbors + compute_pdf of a new neighbour list in planned mode) changes
    how the forward is issued, not one bit of its result."""
    import torch
    from mccnn_b200 import MCConvBuilder, ops
    from mccnn_b200.engine import PlannedForward
    dev = torch.device("cuda:0")
    B = 8
    fn, store = _make(B, 32, 40, dev)
    pts, bids, feats = make_batch(B, 1024, "sphere", seed=21)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    fn(tp, tb, tf)
    _randomize_store(store)
    ref = fn(tp, tb, tf)[:B].clone()
    monkeypatch.setattr(MCConvBuilder, "FUSE_NEIGHBORS_PDF", True)
    calls = []
    orig = ops.find_neighbors_pdf_fused
    monkeypatch.setattr(ops, "find_neighbors_pdf_fused", lambda *a, **kw: (calls.append(1), orig(*a, **kw))[1])
    eng = PlannedForward(fn, B, use_graph=True)
    assert torch.equal(eng(tp, tb, tf), ref)          # records (exact ops)
    assert torch.equal(eng(tp, tb, tf), ref)          # planned: captures the fused launches
    assert torch.equal(eng(tp, tb, tf), ref)          # replay
    assert len(calls) >= 3 and len(calls) % 3 == 0    # MCClassS builds three neighbour lists with densities per run


def test_planned_ops_bit_identical():
    """Hierarchy + Conv_1 (no dense glue in between): planned, capacity-padded results equal the exact ones bit for
    bit on their live rows, for every tensor the builder exposes."""
    import torch
    from mccnn_b200 import runtime, variables
    from mccnn_b200.MCConvBuilder import PointHierarchy, ConvolutionBuilder
    dev = torch.device("cuda:0")
    B, N = 8, 1024
    pts, bids, feats = make_batch(B, N, "sphere", seed=13)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    store = variables.VariableStore(device=dev, seed=5)

    def run():
        ph = PointHierarchy(tp, tf, tb, [0.1, 0.4, 1.9], "H", B)
        cb = ConvolutionBuilder(KDEWindow=0.2, variableStore=store)
        out = cb.create_convolution("c1", ph, 0, tf, 1, 0.2, outPointLevel=1, multiFeatureConv=True,
                                    outNumFeatures=32)
        grid = cb.cacheGrids_[list(cb.cacheGrids_)[0]]
        neigh = cb.cacheNeighs_[list(cb.cacheNeighs_)[0]]
        pdf = cb.cachePDFs_[list(cb.cachePDFs_)[0]]
        return ph, grid, neigh, pdf, out

    with runtime.record() as plan:
        e_ph, e_grid, e_neigh, e_pdf, e_out = run()
    with runtime.planned(plan):
        p_ph, p_grid, p_neigh, p_pdf, p_out = run()
    assert int(plan.overflow.item()) == 0

    def live(t):
        c = runtime.count_of(t)
        return t if c is None else t[: int(c.item())]

    for lvl in range(4):
        assert torch.equal(live(p_ph.points_[lvl]), e_ph.points_[lvl])
        assert torch.equal(live(p_ph.batchIds_[lvl]), e_ph.batchIds_[lvl])
        assert torch.equal(live(p_ph.features_[lvl]), e_ph.features_[lvl])
    for a, b in zip(p_ph.sampledIndexs_, e_ph.sampledIndexs_):
        assert torch.equal(live(a), b)
    for a, b in zip(p_grid, e_grid):
        assert torch.equal(live(a), b)
    for a, b in zip(p_neigh, e_neigh):
        assert torch.equal(live(a), b)
    assert torch.equal(live(p_pdf), e_pdf)
    assert torch.equal(live(p_out), e_out)


def test_capacity_overflow_is_detected_and_repaired():
    import torch
    from mccnn_b200.engine import PlannedForward
    dev = torch.device("cuda:0")
    B, N, k = 4, 1024, 16
    fn, store = _make(B, k, 40, dev)
    sparse = make_batch(B, N, "sphere", seed=21)
    # same shapes, but every cloud collapsed into a tiny ball around a few far-apart centres: far more neighbours
    dense_pts = sparse[0].copy()
    dense_pts[:] = dense_pts * 0.02 + np.sign(dense_pts) * 0.4
    eng = PlannedForward(fn, B, use_graph=True, slack=1.0)
    ts = [torch.from_numpy(a).to(dev) for a in sparse]
    eng(*ts)
    eng(*ts)
    td = [torch.from_numpy(dense_pts).to(dev), ts[1], ts[2]]
    want = fn(*td)[:B].clone()
    got = eng(*td).clone()
    assert torch.equal(got, want)


def test_pipelined_and_host_paths_match_exact():
    """Several forwards in flight (PipelinedForward), packed single-copy inputs and the host-facing calls all return
    the exact-mode logits bit for bit."""
    import torch
    from mccnn_b200.engine import (PlannedForward, PipelinedForward, HostForward, PipelinedHostForward, pack_batch)
    dev = torch.device("cuda:0")
    B, N, k = 4, 1024, 16
    fn, store = _make(B, k, 40, dev)
    host = [make_batch(B, N, "sphere", seed=40 + i) for i in range(5)]
    devb = [tuple(torch.from_numpy(a).to(dev) for a in h) for h in host]
    fn(*devb[0])
    _randomize_store(store)
    want = [fn(*b)[:B].clone() for b in devb]

    pipe = PipelinedForward(fn, B, depth=3)
    pipe.calibrate(devb)
    tickets = []
    got = [None] * len(devb)
    for i, b in enumerate(devb):                       # more batches than lanes: lanes are reused
        if i % 2 == 0:
            t = pipe.submit(*b)
        else:
            t = pipe.submit_packed(pack_batch(*b))
        out, flag = pipe.result(t)
        got[i] = out.clone()
        assert int(flag.item()) == 0
    for g, w in zip(got, want):
        assert torch.equal(g, w)

    eng = PlannedForward(fn, B)
    eng.calibrate(devb)
    hf = HostForward(eng, dev)
    cpu = [tuple(torch.from_numpy(a) for a in h) for h in host]
    for _ in range(2):                                 # first call captures, second replays from pinned staging
        for c, w in zip(cpu, want):
            assert torch.equal(hf(*c), w.cpu())
    blocks = [pack_batch(*c, pin=True) for c in cpu]
    for blk, w in zip(blocks, want):
        assert torch.equal(hf.call_packed(blk), w.cpu())

    phf = PipelinedHostForward(pipe, dev)
    tk = [phf.submit_packed(blk) for blk in blocks[:3]]
    for t, w in zip(tk, want[:3]):
        assert torch.equal(phf.result(t), w.cpu())


def test_mcclasss_matches_oracle_pipeline():
    import torch
    from oracle import pipeline as OP
    dev = torch.device("cuda:0")
    B, N, k = 4, 1024, 16
    fn, store = _make(B, k, 40, dev)
    pts, bids, feats = make_batch(B, N, "sphere", seed=31)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    fn(tp, tb, tf)
    _randomize_store(store)
    got = fn(tp, tb, tf)[:B].cpu().numpy()
    params = {n: v.cpu().numpy() for n, v in store.vars.items()}
    want, trace = OP.mcclasss_forward(pts, bids, feats, 1, B, k, 40, params)
    assert got.shape == want.shape
    denom = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-4 * denom, np.abs(got - want).max() / denom


def test_peer_gather_single_rank_group():
    """engine.PeerGather (csrc/peer_gather.cu, the NVLink peer-memory all-gather that closes a sharded step) on a
    one-rank process group: the same kernel, symmetric-memory rendezvous and CUDA-graph capture as on N GPUs, with the
    rank as its own only peer.  (The N > 1 exchange is exercised by tools/peer_gather_check.py under torchrun and by
    bench.py --gpus N; the driver's GPU test tier has one GPU.)"""
    import os
    import torch
    import torch.distributed as dist
    from mccnn_b200.engine import PeerGather
    dev = torch.device("cuda:0")
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        try:
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        except Exception as exc:  # noqa: BLE001
            pytest.skip("no process group on this box: %r" % (exc,))
        created = True
    try:
        try:
            pg = PeerGather(8, 40, dev)
        except Exception as exc:  # noqa: BLE001
            pytest.skip("symmetric memory unavailable on this box: %r" % (exc,))
        x = torch.arange(320, dtype=torch.float32, device=dev).reshape(8, 40)
        for step in range(5):
            assert torch.equal(pg(x + step), x + step)
        xs = x.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pg(xs)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = pg(xs)
        for step in range(5):
            xs.fill_(float(step))
            g.replay()
            assert torch.equal(out, xs)
        # the put / collect ring (engine.PeerExchange): equal rows, then unequal live rows compacted, with the overflow
        # flag travelling in the header; put captured in a graph, collect launched behind it
        from mccnn_b200.engine import PeerExchange
        px = PeerExchange(8, 40, dev)
        for step in range(9):   # more steps than ring slots
            assert torch.equal(px(x + step), x + step)
        pc = PeerExchange(100, 7, dev, compact=True)
        flag = torch.zeros((1,), dtype=torch.int32, device=dev)
        for n in (1, 37, 100, 64):
            rows = torch.randn((n, 7), device=dev)
            flag.fill_(n == 37)
            pc.put(rows, flag=flag)
            out, meta = pc.collect()
            m = meta.cpu().tolist()
            assert m == [n, int(n == 37), n, int(n == 37)] and torch.equal(out[:n], rows)
        g2 = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            px.put(xs)
            px.collect()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g2):
            px.put(xs)
        for step in range(9):
            xs.fill_(float(step))
            g2.replay()
            assert torch.equal(px.collect()[0], xs)
        # deferred collect (pending_only): a no-op when nothing is outstanding, otherwise the previous step's rows --
        # captured in front of the put, so that a lane's graph never waits on a peer's put of the SAME step
        pd = PeerExchange(8, 40, dev)
        before = pd.out.clone()
        pd.collect(pending_only=True)
        assert torch.equal(pd.out, before) and int(pd.state[1]) == 0        # nothing was outstanding
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pd.collect(pending_only=True)
            pd.put(xs)
        torch.cuda.current_stream().wait_stream(side)
        g3 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g3):
            pd.collect(pending_only=True)
            pd.put(xs)
        prev = xs.clone()                                                     # rows of the warm-up put
        for step in range(9):
            xs.fill_(float(100 + step))
            g3.replay()
            assert torch.equal(pd.out, prev)                                  # step k's graph delivers step k - 1
            prev = xs.clone()
        pd.collect(pending_only=True)                                         # flush
        assert torch.equal(pd.out, prev)
        pd.collect(pending_only=True)
        assert int(pd.state[0]) == int(pd.state[1]) == 10
        # the same through the engine: PipelinedForward(collect_deferred=True) -- join() flushes every lane
        from mccnn_b200.engine import PipelinedForward
        B = 4
        fn, store = _make(B, 16, 40, dev)
        pts, bids, feats = make_batch(B, 512, "sphere", seed=41)
        batch = tuple(torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
        want = fn(*batch)[:B].clone()
        lanes = [PeerExchange(B, 40, dev) for _ in range(2)]
        pipe = PipelinedForward(fn, B, depth=2, exchanges=lanes, collect_deferred=True)
        pipe.calibrate([batch])
        for i in range(5):
            pipe.submit(*batch)
        pipe.join()
        torch.cuda.synchronize()
        for ex in lanes:
            assert int(ex.state[0]) == int(ex.state[1])                       # every put was collected
            assert torch.equal(ex.out, want)
    finally:
        if created:
            dist.destroy_process_group()
