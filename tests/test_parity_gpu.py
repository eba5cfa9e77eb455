"""GPU parity: every op of libmccnn_b200.so, called through the reference-facing Python API (MCConvModule ->
ctypes -> C ABI), against (a) the CPU oracle (oracle/mccnn_oracle.c) and (b) the reference's own CUDA kernels
(oracle/_ref/libmccnn_ref.so) on identical seeded inputs.

Bars (BASELINE.json north_star): bit-exact for aabb / sort / find_neighbors / poisson_sampling indices; 1e-5 relative
for compute_pdf / spatial_conv (float32; the reference's own result is accumulation-order dependent).
Where the reference is order-nondeterministic (within-cell sort order, poisson output order inside a (cloud, phase)
launch) the comparison is on the canonicalised form -- see oracle/mccnn_oracle.c header.
"""
import numpy as np
import pytest

from common import make_batch, make_cloud, make_conv_weights, rel_err, sum_err

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for compute_pdf / spatial_conv


@pytest.fixture(scope="module")
def env():
    import torch
    import mccnn_b200.MCConvModule as M
    from oracle import cpu as O
    from oracle import ref as R
    dev = torch.device("cuda:0")

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def A(t):
        return t.detach().cpu().numpy()

    return dict(torch=torch, M=M, O=O, R=R if R.available() else None, T=T, A=A, dev=dev)


CASES = [
    # (batch, points per cloud, distribution, ragged, grid radius, scale_inv)
    (1, 1024, "cube", False, 0.1, True),
    (1, 1024, "sphere", False, 0.1, True),
    (4, 1024, "sphere", True, 0.2, True),
    (32, 1024, "sphere", False, 0.1, True),
    (3, 700, "cube", True, 0.4, True),
    (2, 333, "sphere", False, 1.9, True),     # n = 1: whole cloud in one cell
    (2, 5000, "room", False, 0.25, False),    # absolute radius, batch-global box (ScanNet style)
    (2, 2000, "lattice", True, 0.25, False),  # duplicates, points on cell boundaries, distances == r exactly
    (3, 1500, "lattice", False, 0.125, True), # the same with per-cloud relative radius (extent 2 -> r = 0.25)
]


def _grid(e, pts, bids, feats, B, radius, scaleInv, impl):
    mn, mx = impl.compute_aabb(pts, bids, B, scaleInv)
    keys, idx = impl.sort_points_step1(pts, bids, mn, mx, B, radius, scaleInv)
    sp, sb, sf, cells = impl.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, radius, scaleInv)
    return mn, mx, keys, idx, sp, sb, sf, cells


@pytest.mark.parametrize("B,N,dist,ragged,radius,scaleInv", CASES)
def test_aabb_sort_bit_exact(env, B, N, dist, ragged, radius, scaleInv):
    e = env
    pts, bids, feats = make_batch(B, N, dist, seed=1, num_features=3, ragged=ragged)
    o = _grid(e, pts, bids, feats, B, radius, scaleInv, e["O"])
    g = _grid(e, e["T"](pts), e["T"](bids), e["T"](feats), B, radius, scaleInv, e["M"])
    names = ["aabbMin", "aabbMax", "keys", "indexs", "sortPts", "sortBatchIds", "sortFeatures", "cellIndexs"]
    for name, a, b in zip(names, g, o):
        assert np.array_equal(e["A"](a).reshape(-1), np.asarray(b).reshape(-1)), name
    # the fused one-launch grid build (builder path) gives the same six tensors
    fk, fi, fp, fb, ff, fc = e["M"].sort_points(e["T"](pts), e["T"](bids), e["T"](feats), g[0], g[1], B, radius,
                                                 scaleInv)
    for name, a, b in zip(names[2:], (fk, fi, fp, fb, ff, fc), o[2:]):
        assert np.array_equal(e["A"](a).reshape(-1), np.asarray(b).reshape(-1)), "fused " + name
    tail = pts.shape[0] % 256  # reference aabb leaves smem rows uninitialised when 0 < N % 256 < B (aabb_gpu.cu:69-78)
    if e["R"] is not None and (tail == 0 or tail >= B):
        R, T, A = e["R"], e["T"], e["A"]
        tb = T(bids)
        rmn, rmx = R.compute_aabb(T(pts), tb, B, scaleInv)
        assert np.array_equal(A(rmn), A(g[0])) and np.array_equal(A(rmx), A(g[1]))
    if e["R"] is not None:
        R, T, A = e["R"], e["T"], e["A"]
        rkeys, ridx = R.sort_points_step1(T(pts), T(bids), g[0], g[1], B, radius, scaleInv)
        assert np.array_equal(A(rkeys), A(g[2]))
        ridx = A(ridx)
        assert np.array_equal(np.sort(ridx), np.arange(pts.shape[0]))          # a permutation ...
        skeys = np.empty_like(A(rkeys))
        skeys[ridx] = A(rkeys)
        assert np.all(np.diff(skeys) >= 0)                                      # ... that sorts by key
        rp, rb, rf, rcells = R.sort_points_step2(T(pts), T(bids), T(feats), rkeys, T(ridx.astype(np.int32)), g[0],
                                                 g[1], B, radius, scaleInv)
        assert np.array_equal(A(rcells), A(g[7]))                               # cell table is order independent


@pytest.mark.parametrize("B,N,dist,ragged,radius,scaleInv", CASES)
def test_find_neighbors_bit_exact(env, B, N, dist, ragged, radius, scaleInv):
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    pts, bids, feats = make_batch(B, N, dist, seed=2, num_features=1, ragged=ragged)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    # samples = a Poisson sub-sampling (different set than the points) and the points themselves
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    for samples, sbids in ((smp, smpb), (sp, sb)):
        ostart, opairs = O.find_neighbors(samples, sbids, sp, cells, mn, mx, radius, B, scaleInv)
        gstart, gpairs = M.find_neighbors(T(samples), T(sbids), T(sp), T(cells), T(mn), T(mx), radius, B, scaleInv)
        assert np.array_equal(A(gstart), ostart)
        assert np.array_equal(A(gpairs), opairs)
        if e["R"] is not None:
            rstart, rpairs = e["R"].find_neighbors(T(samples), T(sbids), T(sp), T(cells), T(mn), T(mx), radius, B,
                                                   scaleInv)
            assert np.array_equal(A(rstart), A(gstart))
            assert np.array_equal(A(rpairs), A(gpairs))


@pytest.mark.parametrize("B,N,dist,ragged,radius,scaleInv", CASES + [(4, 4000, "sphere", False, 0.03, True),
                                                                      (2, 20000, "room", False, 0.1, False)])
def test_find_neighbors_fused_bit_exact(env, B, N, dist, ragged, radius, scaleInv):
    """The one-launch count + scan + fill kernel of the sync-free pipeline (mc_find_neighbors_fused: candidate-per-lane
    sweep for populated cells, lane-per-cell sweep for sparse grids, decoupled look-back for the start indices) gives
    the oracle's start indices and pair list bit for bit; a too small capacity raises the overflow flag."""
    e = env
    O, T, A, torch = e["O"], e["T"], e["A"], e["torch"]
    from mccnn_b200 import ops
    pts, bids, feats = make_batch(B, N, dist, seed=12, num_features=1, ragged=ragged)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    for samples, sbids in ((smp, smpb), (sp, sb)):
        ostart, opairs = O.find_neighbors(samples, sbids, sp, cells, mn, mx, radius, B, scaleInv)
        K = opairs.shape[0]
        flag = torch.zeros((1,), dtype=torch.int32, device=e["dev"])
        start, pairs, k_dev = ops.find_neighbors_fused(T(samples), T(sbids).reshape(-1), T(sp), T(cells), T(mn), T(mx),
                                                       radius, scaleInv, K + 64, overflow=flag)
        assert int(k_dev.item()) == K and int(flag.item()) == 0
        assert np.array_equal(A(start), ostart.reshape(-1))
        assert np.array_equal(A(pairs)[:K], opairs)
        if K > 8:
            flag.zero_()
            ops.find_neighbors_fused(T(samples), T(sbids).reshape(-1), T(sp), T(cells), T(mn), T(mx), radius, scaleInv,
                                     K - 5, overflow=flag)
            assert int(flag.item()) == 1


@pytest.mark.parametrize("B,N,dist,ragged,radius,scaleInv", CASES + [(4, 4000, "sphere", False, 0.03, True),
                                                                      (2, 20000, "room", False, 0.1, False),
                                                                      (1, 3000, "cube", False, 0.6, False),
                                                                      (8, 4096, "sphere", False, 0.1, False)])
def test_find_neighbors_pdf_fused_equals_the_two_ops(env, B, N, dist, ragged, radius, scaleInv):
    """mc_find_neighbors_pdf_fused (neighbour list + densities in one launch) == mc_find_neighbors_fused followed by
    mc_compute_pdf: same start indices and pairs, and the SAME float for every density (same terms in the same order),
    including neighbour lists longer than the kernel's per-warp coordinate stage (the radius-0.6 case) and calls above
    the sample count at which the entry point hands the densities to the pair-per-thread kernel (8 x 4096)."""
    e = env
    O, T, A, torch = e["O"], e["T"], e["A"], e["torch"]
    from mccnn_b200 import ops
    pts, bids, feats = make_batch(B, N, dist, seed=14, num_features=1, ragged=ragged)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    for samples, sbids in ((smp, smpb), (sp, sb)):
        ostart, opairs = O.find_neighbors(samples, sbids, sp, cells, mn, mx, radius, B, scaleInv)
        K = opairs.shape[0]
        flag = torch.zeros((1,), dtype=torch.int32, device=e["dev"])
        geo = (T(samples), T(sbids).reshape(-1), T(sp), T(cells), T(mn), T(mx))
        start, pairs, pdf, k_dev = ops.find_neighbors_pdf_fused(*geo, radius, 0.25, scaleInv, K + 64, overflow=flag)
        assert int(k_dev.item()) == K and int(flag.item()) == 0
        assert np.array_equal(A(start), ostart.reshape(-1))
        assert np.array_equal(A(pairs)[:K], opairs)
        want = ops.compute_pdf(T(sp), T(sb).reshape(-1), T(mn), T(mx), start, pairs[:K].contiguous(), 0.25, radius,
                               scaleInv)
        assert torch.equal(pdf[:K], want.reshape(-1))
        opdf = O.compute_pdf(sp, sb, mn, mx, ostart, opairs, 0.25, radius, B, scaleInv)
        np.testing.assert_allclose(A(pdf)[:K], opdf.reshape(-1), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("B,N,dist,ragged,radius,scaleInv", CASES)
def test_poisson_bit_exact(env, B, N, dist, ragged, radius, scaleInv):
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    pts, bids, feats = make_batch(B, N, dist, seed=3, num_features=2, ragged=ragged)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    op, ob, oi, oph = O.poisson_sampling(sp, sb, cells, mn, mx, radius, B, scaleInv, return_phase=True)
    gp, gb, gi = M.poisson_sampling(T(sp), T(sb), T(cells), T(mn), T(mx), radius, B, scaleInv)
    assert np.array_equal(A(gi), oi)
    assert np.array_equal(A(gb), ob)
    assert np.array_equal(A(gp), op)
    # gathers that follow in the hierarchy
    gf = M.get_sampled_features(gi, T(sf))
    assert np.array_equal(A(gf), O.get_sampled_features(oi, sf))
    gt = M.transform_indexs(gi, T(idx))
    assert np.array_equal(A(gt), O.transform_indexs(oi, idx))
    if e["R"] is not None:
        rp, rb, ri = e["R"].poisson_sampling(T(sp), T(sb), T(cells), T(mn), T(mx), radius, B, scaleInv)
        ri, rb = A(ri), A(rb).reshape(-1)
        assert ri.shape == oi.shape
        assert np.array_equal(np.sort(ri), np.sort(oi))               # identical accepted set
        # the reference emits (cloud, phase) groups in order; inside a group the order is arrival order
        phase_of = dict(zip(oi.tolist(), oph.tolist()))
        group = np.array([rb[k] * 27 + phase_of[int(ri[k])] for k in range(ri.shape[0])])
        assert np.all(np.diff(group) >= 0)
        assert np.array_equal(A(rp), sp[ri])


@pytest.mark.parametrize("sizes,radius", [((3000, 100, 100, 100), 0.1),       # one cloud above the staged-points capacity
                                          ((6000,) + (50,) * 11, 0.08),        # ... and above the shared-memory flags
                                          ((40, 5000, 40, 0, 700, 40), 0.12)])  # empty cloud in the middle
def test_poisson_bit_exact_on_very_ragged_batches(env, sizes, radius):
    """poisson_kernel sizes its shared-memory staging for twice the MEAN cloud of the batch (choose_caps); a cloud far
    above the mean takes the global-memory paths of the same kernel and must give the same samples."""
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    B = len(sizes)
    pts = np.concatenate([make_cloud(n, "sphere", np.random.default_rng(70 + b)).reshape(-1, 3)
                          for b, n in enumerate(sizes)]).astype(np.float32)
    bids = np.concatenate([np.full((n, 1), b, np.int32) for b, n in enumerate(sizes)])
    feats = np.ones((pts.shape[0], 1), np.float32)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, False, O)
    op, ob, oi = O.poisson_sampling(sp, sb, cells, mn, mx, radius, B, False)
    gp, gb, gi = M.poisson_sampling(T(sp), T(sb), T(cells), T(mn), T(mx), radius, B, False)
    assert np.array_equal(A(gi), oi)
    assert np.array_equal(A(gb), ob)
    assert np.array_equal(A(gp), op)


PDF_CASES = [(1, 1024, "cube", 0.1, 0.2, True), (1, 1024, "sphere", 0.1, 0.2, True),
             (8, 1024, "sphere", 0.2, 0.2, True), (4, 600, "sphere", 0.8, 0.25, True),
             (2, 4000, "room", 0.3, 0.25, False)]


@pytest.mark.parametrize("B,N,dist,radius,window,scaleInv", PDF_CASES)
def test_compute_pdf(env, B, N, dist, radius, window, scaleInv):
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    pts, bids, feats = make_batch(B, N, dist, seed=4)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, radius, B, scaleInv)
    opdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, window, radius, B, scaleInv)
    gpdf = M.compute_pdf(T(sp), T(sb), T(mn), T(mx), T(start), T(pairs), window, radius, B, scaleInv)
    assert gpdf.shape == (pairs.shape[0], 1)
    err = rel_err(A(gpdf), opdf, atol_scale=0.0)
    assert err <= RTOL, err
    if e["R"] is not None:
        rpdf = e["R"].compute_pdf(T(sp), T(sb), T(mn), T(mx), T(start), T(pairs), window, radius, B, scaleInv)
        err = rel_err(A(gpdf), A(rpdf), atol_scale=0.0)
        assert err <= RTOL, err


CONV_CASES = [
    # B, N, dist, C, Co, combin, radius, avg, scaleInv
    (1, 1024, "cube", 3, 8, True, 0.1, True, True),      # BASELINE config 1
    (1, 1024, "sphere", 3, 8, True, 0.1, True, True),
    (4, 1024, "sphere", 1, 64, True, 0.2, True, True),    # MCClassS Conv_1
    (4, 400, "sphere", 128, 128, False, 0.8, True, True),  # MCClassS Conv_2 shape (depth-wise)
    (2, 300, "cube", 16, 16, False, 0.5, False, True),     # avg = False
    (2, 500, "cube", 4, 5, True, 0.3, True, True),         # neurons not a multiple of 8, outputs span blocks
    (2, 3000, "room", 4, 32, True, 0.3, True, False),      # absolute radius
]


# both kernels behind mc_spatial_conv are held to every case: the warp-level one (small convolutions), and the
# persistent producer / consumer one in its CTA shapes "consumer warps, blocks per warp" (csrc/spatial_conv.cu picks by
# size; MCCNN_CONV_CFG forces a variant)
@pytest.mark.parametrize("variant", ["auto", "warp", "8,2", "4,4", "4,1", "8,1"])
@pytest.mark.parametrize("B,N,dist,C,Co,combin,radius,avg,scaleInv", CONV_CASES)
def test_spatial_conv(env, monkeypatch, variant, B, N, dist, C, Co, combin, radius, avg, scaleInv):
    if variant != "auto":
        monkeypatch.setenv("MCCNN_CONV_CFG", variant)
    else:
        monkeypatch.delenv("MCCNN_CONV_CFG", raising=False)
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    pts, bids, feats = make_batch(B, N, dist, seed=5, num_features=C)
    if C == 1:
        feats = np.random.default_rng(9).uniform(-1, 1, feats.shape).astype(np.float32)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, scaleInv, O)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, radius, B, scaleInv)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.2, radius, B, scaleInv)
    w1, w2, w3, b1, b2, b3 = make_conv_weights(C, Co, combin, seed=6)
    oout = O.spatial_conv(sp, sf, sb, pdf, smp, start, pairs, mn, mx, w1, w2, w3, b1, b2, b3, Co, combin, B, radius,
                          scaleInv, avg, accum_double=True)
    gout = M.spatial_conv(T(sp), T(sf), T(sb), T(pdf), T(smp), T(start), T(pairs), T(mn), T(mx), T(w1), T(w2), T(w3),
                          T(b1), T(b2), T(b3), Co, combin, B, radius, scaleInv, avg)
    assert tuple(gout.shape) == oout.shape
    # an output is a signed float32 sum over neighbours x input features: its error is measured against the sum of
    # the absolute contributions (any summation order, the reference's atomics included, is only accurate on that
    # scale); elements that do not cancel are additionally held to the plain 1e-5 relative bound.
    scale = O.spatial_conv(sp, sf, sb, pdf, smp, start, pairs, mn, mx, w1, w2, w3, b1, b2, b3, Co, combin, B, radius,
                           scaleInv, avg, abs_sum=True)
    err = sum_err(A(gout), oout, scale)
    assert err <= RTOL, err
    big = np.abs(oout) >= 0.1 * scale
    if big.any():
        assert rel_err(A(gout)[big], oout[big], atol_scale=0.0) <= 10 * RTOL
    if e["R"] is not None:
        rout = e["R"].spatial_conv(T(sp), T(sf), T(sb), T(pdf), T(smp), T(start), T(pairs), T(mn), T(mx), T(w1),
                                   T(w2), T(w3), T(b1), T(b2), T(b3), Co, combin, B, radius, scaleInv, avg)
        err_vs_ref = sum_err(A(gout), A(rout), scale)
        err_ref = sum_err(A(rout), oout, scale)   # the reference's own distance from the order-free value
        assert err_vs_ref <= RTOL, (err_vs_ref, err_ref)
        assert err <= max(2 * err_ref, 2e-6), (err, err_ref)  # and we are not less accurate than the reference


@pytest.mark.parametrize("variant", ["8,2", "4,4", "8,1"])
@pytest.mark.parametrize("C,Co,combin", [(64, 64, False), (4, 32, True)])
def test_persistent_conv_is_bit_reproducible_under_load(env, monkeypatch, variant, C, Co, combin):
    """The persistent kernel hands its shared-memory stages from producer to consumer warps through raw mbarriers
    (full / empty rings).  compute-sanitizer's racecheck does not model those (it tracks cuda::barrier objects), so it
    reports the stage reuse as WAR hazards; a real race on a stage would show up as run-to-run differences.  The same
    convolution is run 24 times, alternating with an L2-thrashing kernel that shifts the relative timing of producers
    and consumers, and every result must equal the first bit for bit (and the warp-level kernel's result, which shares
    no staging code, to the usual tolerance)."""
    e = env
    O, M, T, A, torch = e["O"], e["M"], e["T"], e["A"], e["torch"]
    B, N, radius = 8, 2048, 0.25
    pts, bids, feats = make_batch(B, N, "sphere", seed=31, num_features=C)
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, B, radius, True, O)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, True)
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, radius, B, True)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.2, radius, B, True)
    w = make_conv_weights(C, Co, combin, seed=32)
    args = [T(x) for x in (sp, sf, sb, pdf, smp, start, pairs, mn, mx) + tuple(w)]
    monkeypatch.setenv("MCCNN_CONV_CFG", "warp")
    ref = M.spatial_conv(*args, Co, combin, B, radius, True, True).clone()
    monkeypatch.setenv("MCCNN_CONV_CFG", variant)
    first = M.spatial_conv(*args, Co, combin, B, radius, True, True).clone()
    scale = first.abs().max().item()
    assert float((first - ref).abs().max()) <= 1e-4 * scale
    thrash = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=e["dev"])
    side = torch.cuda.Stream()
    for rep in range(24):
        if rep % 2:
            with torch.cuda.stream(side):   # competes for SMs and L2 while the convolution runs
                thrash.add_(1.0)
        out = M.spatial_conv(*args, Co, combin, B, radius, True, True)
        assert torch.equal(out, first), rep
    torch.cuda.synchronize()


def test_feature_permutations(env):
    e = env
    O, M, T, A = e["O"], e["M"], e["T"], e["A"]
    rng = np.random.default_rng(0)
    for N, C in ((1000, 1), (777, 3), (4096, 64), (513, 128)):
        f = rng.uniform(-1, 1, (N, C)).astype(np.float32)
        perm = rng.permutation(N).astype(np.int32)
        assert np.array_equal(A(M.sort_features(T(f), T(perm))), O.sort_features(f, perm))
        assert np.array_equal(A(M.sort_features_back(T(f), T(perm))), O.sort_features_back(f, perm))
        back = M.sort_features_back(M.sort_features(T(f), T(perm)), T(perm))
        assert np.array_equal(A(back), f)  # round trip


def test_edge_cases(env):
    e = env
    O, M, T, A, torch = e["O"], e["M"], e["T"], e["A"], e["torch"]
    # batch declared larger than the clouds present (eval builds the graph with batchSize 32 and feeds one cloud)
    pts, bids, feats = make_batch(1, 512, "sphere", seed=8)
    gmn, gmx = M.compute_aabb(T(pts), T(bids), 4, True)
    omn, omx = O.compute_aabb(pts, bids, 4, True)
    assert np.array_equal(A(gmn), omn) and np.array_equal(A(gmx), omx)
    # a sample with no neighbour produces a zero output row and an empty CSR row
    mn, mx, keys, idx, sp, sb, sf, cells = _grid(e, pts, bids, feats, 1, 0.05, True, O)
    far = np.array([[0.0, 0.0, 0.0]], np.float32)  # centre of the sphere: nothing within 0.05
    fb = np.zeros((1, 1), np.int32)
    gs, gp = M.find_neighbors(T(far), T(fb), T(sp), T(cells), T(mn), T(mx), 0.05, 1, True)
    assert gp.shape[0] == 0 and A(gs).tolist() == [[0]]
    # a packed batch whose clouds are interleaved (not contiguous): the general path, and the fused path's device-side
    # contiguity check falling back to it, both match the oracle
    p2, b2, f2 = make_batch(3, 400, "cube", seed=9, num_features=2)
    perm = np.random.default_rng(1).permutation(p2.shape[0])
    p2, b2, f2 = p2[perm], b2[perm], f2[perm]
    omn, omx = O.compute_aabb(p2, b2, 3, True)
    gmn, gmx = M.compute_aabb(T(p2), T(b2), 3, True)
    assert np.array_equal(A(gmn), omn) and np.array_equal(A(gmx), omx)
    ok, oi = O.sort_points_step1(p2, b2, omn, omx, 3, 0.2, True)
    osp, osb, osf, ocells = O.sort_points_step2(p2, b2, f2, ok, oi, omn, omx, 3, 0.2, True)
    fused = M.sort_points(T(p2), T(b2), T(f2), gmn, gmx, 3, 0.2, True)
    for a, b in zip(fused, (ok, oi, osp, osb, osf, ocells)):
        assert np.array_equal(A(a).reshape(-1), np.asarray(b).reshape(-1))
    # invalid attributes raise like OP_REQUIRES
    with pytest.raises(ValueError):
        M.find_neighbors(T(far), T(fb), T(sp), T(cells), T(mn), T(mx), -1.0, 1, True)
    with pytest.raises(ValueError):
        M.compute_aabb(T(pts), T(bids), 0, True)
    with pytest.raises(ValueError):
        M.sort_features(T(feats), T(np.zeros((3,), np.int32)))


def test_dense_glue_matches_torch(env):
    """mc_scale_shift_act / mc_dense_layer (csrc/dense.cu) against the plain PyTorch fp32 formulation of the
    reference's TensorFlow glue (utils/MCNetworkUtils.py:20-142), 1e-5 relative."""
    e = env
    torch, T, A = e["torch"], e["T"], e["A"]
    from mccnn_b200 import ops
    rng = np.random.default_rng(3)
    for rows, K, N in ((5400, 64, 128), (387, 128, 256), (32, 256, 128), (32, 64, 40), (7, 5, 3), (100, 33, 70)):
        x = T(rng.normal(size=(rows, K)).astype(np.float32))
        w = T((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32))
        b = T(rng.normal(size=(N,)).astype(np.float32) * 0.1)
        s1, t1 = T(rng.uniform(0.5, 1.5, K).astype(np.float32)), T(rng.normal(size=K).astype(np.float32) * 0.1)
        s2, t2 = T(rng.uniform(0.5, 1.5, N).astype(np.float32)), T(rng.normal(size=N).astype(np.float32) * 0.1)
        want = torch.relu((torch.relu(x.double() * s1.double() + t1.double()) @ w.double() + b.double()) * s2.double()
                          + t2.double())
        got = ops.dense_layer(x, w, b, pre=(s1, t1), pre_relu=True, post=(s2, t2), post_relu=True)
        scale = (torch.relu(x.double() * s1.double() + t1.double()).abs() @ w.double().abs() + b.double().abs()) \
            * s2.double() + t2.double().abs()
        assert float(((got.double() - want).abs() / scale).max()) <= RTOL
        plain = ops.dense_layer(x, w)
        ref = x.double() @ w.double()
        assert float(((plain.double() - ref).abs() / (x.double().abs() @ w.double().abs() + 1e-30)).max()) <= RTOL
        y = ops.scale_shift_act(x, s1, t1, True)
        assert float((y - torch.relu(torch.addcmul(t1, x, s1))).abs().max()) <= 1e-6
    # live row count: rows past the count are not touched
    x = T(rng.normal(size=(64, 16)).astype(np.float32))
    w = T(rng.normal(size=(16, 8)).astype(np.float32))
    cnt = torch.tensor([10], dtype=torch.int32, device=e["dev"])
    full = ops.dense_layer(x, w)
    part = ops.dense_layer(x, w, n_dev=cnt)
    assert torch.equal(part[:10], full[:10])


def test_dense_fewrows_matches_torch(env):
    """mc_dense_layer_fewrows (few rows x wide layer: MCClassH's 2048 -> 1024 -> 512 -> 40 head) against float64
    torch, 1e-5 relative to the sum of |terms|; capacity rows past the live count are not touched; a second call gives
    the same bits (the K segments are summed in a fixed order)."""
    e = env
    torch, T = e["torch"], e["T"]
    from mccnn_b200 import ops
    rng = np.random.default_rng(5)
    for rows, K, N in ((32, 2048, 1024), (32, 1024, 512), (32, 512, 40), (16, 2048, 1024), (1, 300, 70), (97, 640, 257),
                       (8, 64, 8), (128, 1664, 64)):
        x = T(rng.normal(size=(rows, K)).astype(np.float32))
        w = T((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32))
        b = T(rng.normal(size=(N,)).astype(np.float32) * 0.1)
        s1, t1 = T(rng.uniform(0.5, 1.5, K).astype(np.float32)), T(rng.normal(size=K).astype(np.float32) * 0.1)
        s2, t2 = T(rng.uniform(0.5, 1.5, N).astype(np.float32)), T(rng.normal(size=N).astype(np.float32) * 0.1)
        a = torch.relu(x.double() * s1.double() + t1.double())
        want = torch.relu((a @ w.double() + b.double()) * s2.double() + t2.double())
        scale = (a.abs() @ w.double().abs() + b.double().abs()) * s2.double() + t2.double().abs()
        got = ops.dense_layer_fewrows(x, w, b, pre=(s1, t1), pre_relu=True, post=(s2, t2), post_relu=True)
        assert float(((got.double() - want).abs() / scale).max()) <= RTOL, (rows, K, N)
        assert torch.equal(got, ops.dense_layer_fewrows(x, w, b, pre=(s1, t1), pre_relu=True, post=(s2, t2),
                                                        post_relu=True))
        plain = ops.dense_layer_fewrows(x, w)
        ref = x.double() @ w.double()
        assert float(((plain.double() - ref).abs() / (x.double().abs() @ w.double().abs() + 1e-30)).max()) <= RTOL
        # capacity-padded input: only the live rows are computed, with the same bits
        live = max(1, rows // 2)
        cnt = torch.tensor([live], dtype=torch.int32, device=e["dev"])
        xp = torch.cat([x, torch.full((160, K), float("nan"), device=e["dev"])])
        part = ops.dense_layer_fewrows(xp, w, b, pre=(s1, t1), pre_relu=True, post=(s2, t2), post_relu=True, n_dev=cnt)
        assert torch.equal(part[:live], got[:live])


def test_deferred_batch_norm_is_folded_exactly(env):
    """An inference batch norm (+ ReLU) is returned as a pending affine (runtime.DeferredAffine) and folded into its
    consumer.  The fused scatter (mc_sort_features_affine) equals scale_shift_act followed by sort_features bit for
    bit; the dense prologue equals the materialised path bit for bit; anything else sees an ordinary tensor."""
    e = env
    torch, T = e["torch"], e["T"]
    from mccnn_b200 import ops, runtime, variables
    from mccnn_b200 import MCConvModule as M, MCNetworkUtils as U
    rng = np.random.default_rng(9)
    for rows, C in ((5400, 128), (387, 256), (33, 7), (1, 4)):
        x = T(rng.normal(size=(rows, C)).astype(np.float32))
        s, t = T(rng.uniform(0.5, 1.5, C).astype(np.float32)), T(rng.normal(size=C).astype(np.float32) * 0.1)
        perm = torch.from_numpy(rng.permutation(rows).astype(np.int32)).to(e["dev"])
        for relu in (True, False):
            want = ops.sort_features(ops.scale_shift_act(x, s, t, relu), perm)
            got = M.sort_features(runtime.DeferredAffine(x, s, t, relu), perm)
            assert torch.equal(got, want)
            assert float((want[perm.long()] - (torch.relu(torch.addcmul(t, x, s)) if relu
                                               else torch.addcmul(t, x, s))).abs().max()) <= 1e-6
    store = variables.VariableStore(device=e["dev"], seed=1)
    x = T(rng.normal(size=(300, 64)).astype(np.float32))
    d = U.batch_norm_RELU_drop_out("bn", x, False, False, 1.0, store)
    assert isinstance(d, runtime.DeferredAffine) and tuple(d.shape) == (300, 64)
    for n in ("bn_BN/gamma", "bn_BN/beta", "bn_BN/moving_mean", "bn_BN/moving_variance"):
        store.vars[n].copy_(T(rng.uniform(0.5, 1.5, 64).astype(np.float32)))
    d = U.batch_norm_RELU_drop_out("bn", x, False, False, 1.0, store)
    fused = U.conv_1x1("c", d, 64, 96, store)                       # prologue of the dense layer
    plain = U.conv_1x1("c", d.materialize(), 64, 96, store)         # separate launch
    assert torch.equal(fused, plain)
    g, b = store.vars["bn_BN/gamma"], store.vars["bn_BN/beta"]
    mu, var = store.vars["bn_BN/moving_mean"], store.vars["bn_BN/moving_variance"]
    want = torch.relu((x - mu) * torch.rsqrt(var + 1e-3) * g + b)
    assert float((torch.add(d, 0.0) - want).abs().max()) <= 1e-5    # torch function on the deferred object
    assert float((d[:5] - want[:5]).abs().max()) <= 1e-5 and d.sum().isfinite()
    # a second affine in front of a dense layer (the initial batch norm of MLP_2_hidden) is folded into the weights
    out = U.MLP_2_hidden(d, 64, 32, 16, 10, "head", 1.0, False, store=store)
    ref = U.MLP_2_hidden(d.materialize(), 64, 32, 16, 10, "head", 1.0, False, store=store)
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_mlp_rows_equals_layer_by_layer(env):
    """mc_mlp_rows (a whole classification head in one launch) does the same arithmetic in the same order as one
    mc_dense_layer launch per layer: bit-identical, and 1e-5 relative against the float64 formulation."""
    e = env
    torch, T = e["torch"], e["T"]
    from mccnn_b200 import ops
    rng = np.random.default_rng(17)
    for rows, dims in ((32, (256, 128, 64, 40)), (5, (7, 300, 3)), (64, (1024, 16)), (1, (3, 5, 7, 9, 11))):
        x = T(rng.normal(size=(rows, dims[0])).astype(np.float32))
        pre = (T(rng.uniform(0.5, 1.5, dims[0]).astype(np.float32)), T(rng.normal(size=dims[0]).astype(np.float32)))
        layers = []
        for k, n in zip(dims[:-1], dims[1:]):
            w = T((rng.normal(size=(k, n)) / np.sqrt(k)).astype(np.float32))
            b = T(rng.normal(size=n).astype(np.float32) * 0.1)
            post = (T(rng.uniform(0.5, 1.5, n).astype(np.float32)), T(rng.normal(size=n).astype(np.float32) * 0.1))
            layers.append((w, b, post, True))
        layers[-1] = (layers[-1][0], None, None, False)
        got = ops.mlp_rows(x, layers, pre=pre, pre_relu=True)
        want = x
        ref = torch.relu(x.double() * pre[0].double() + pre[1].double())
        for i, (w, b, post, relu) in enumerate(layers):
            want = ops.dense_layer(want, w, b, pre=pre if i == 0 else None, pre_relu=(i == 0), post=post,
                                   post_relu=relu, rows_hint=1)
            ref = ref @ w.double()
            if b is not None:
                ref = ref + b.double()
            if post is not None:
                ref = ref * post[0].double() + post[1].double()
            if relu:
                ref = torch.relu(ref)
        assert torch.equal(got, want)
        assert float((got.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    cnt = torch.tensor([3], dtype=torch.int32, device=e["dev"])
    x = T(rng.normal(size=(8, 16)).astype(np.float32))
    w = T(rng.normal(size=(16, 4)).astype(np.float32))
    part = ops.mlp_rows(x, [(w, None, None, False)], n_dev=cnt)
    assert torch.equal(part[:3], ops.mlp_rows(x, [(w, None, None, False)])[:3])
