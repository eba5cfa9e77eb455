"""Backward ops on the GPU through the C ABI: SpatialConvGrad, SortPointsStep2Grad,
GetSampledFeaturesGrad and the autograd wiring of tf_ops/MCConvModuleSrc:28-81, against
  * the CPU oracle (float64 accumulation; itself pinned by torch autograd in tests/test_oracle_grad_cpu.py), and
  * the reference's own gradient kernels (oracle/_ref: spatialConvGradsCPU & co compiled unmodified) on the same inputs.
Tolerance: 1e-5 of each sum's |terms| scale (floored by the tensor's mean magnitude), as for the forward op; the
reference accumulates with float atomics in arbitrary order, ours in a fixed order (parameter gradients are
bit-reproducible run to run)."""
import numpy as np
import pytest

from common import make_batch, make_conv_weights

pytestmark = pytest.mark.gpu
NAMES = ["features", "weights1", "biases1", "weights2", "biases2", "weightsOut", "biasesOut"]


def _case(B, n, dist, r, C, numOut, combin, seed, stride):
    from oracle import cpu as O
    pts, bids, feats = make_batch(B, n, dist, seed=seed, num_features=C, ragged=True)
    mn, mx = O.compute_aabb(pts, bids, B, True)
    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, r, True)
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, r, True)
    smp, smpb = sp[::stride].copy(), sb[::stride].copy()
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, r, B, True)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.25, r, B, True)
    w = make_conv_weights(C, numOut, combin, seed=seed)
    outC = numOut if combin else C
    G = np.random.default_rng(seed).uniform(-1, 1, (smp.shape[0], outC)).astype(np.float32)
    # OP order: W1, b1, W2, b2, Wout, bout
    return dict(geo=(sp, sf, sb, pdf, smp, start, pairs, mn, mx), w=(w[0], w[3], w[1], w[4], w[2], w[5]), G=G,
                attrs=(numOut, combin, B, r, True))


def _err(got, want, scale):
    got, want, scale = (np.asarray(a, np.float64) for a in (got, want, scale))
    return float((np.abs(got - want) / (np.maximum(scale, np.abs(want).mean()) + 1e-30)).max())


@pytest.mark.parametrize("C,numOut,combin,avg,n,stride", [
    (3, 8, True, True, 1024, 1),      # BASELINE config 1 shape
    (1, 64, True, True, 1024, 4),     # MCClassS Conv_1 shape: one input feature, pooled samples
    (16, 16, False, True, 512, 3),    # depth-wise
    (64, 64, False, False, 256, 2),   # depth-wise, no averaging
    (2, 5, True, False, 300, 1),      # partial last block (10 neurons)
    (4, 3, True, True, 700, 5)])
def test_spatial_conv_grad_matches_oracle_and_reference(C, numOut, combin, avg, n, stride):
    import torch
    import mccnn_b200.MCConvModule as M
    from oracle import cpu as O, ref as R
    dev = torch.device("cuda:0")
    cs = _case(2, n, "sphere", 0.2, C, numOut, combin, seed=C * 7 + numOut, stride=stride)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    geo, w, G, attrs = cs["geo"], cs["w"], cs["G"], cs["attrs"] + (avg,)
    want = O.spatial_conv_grad(*geo, *w, G, *attrs)
    scale = O.spatial_conv_grad(*geo, *w, G, *attrs, abs_sum=True)
    tg, tw, tG = [T(a) for a in geo], [T(a) for a in w], T(G)
    got = M.spatial_conv_grad(*tg, *tw, tG, *attrs)
    again = M.spatial_conv_grad(*tg, *tw, tG, *attrs)
    for name, g, g2, wnt, sc in zip(NAMES, got, again, want, scale):
        assert tuple(g.shape) == wnt.shape, name
        assert _err(g.cpu().numpy(), wnt, sc) <= 1e-5, name
        if name != "features":  # parameter gradients: fixed summation order
            assert torch.equal(g, g2), name
    if R.available():
        ref = R.spatial_conv_grad(*tg, *tw, tG, *attrs)
        R.lib().ref_sync()
        for name, g, rf, sc in zip(NAMES, got, ref, scale):
            assert _err(g.cpu().numpy(), rf.cpu().numpy(), sc) <= 2e-5, name


def test_permutation_grads_match_oracle_and_reference():
    import torch
    import mccnn_b200.MCConvModule as M
    from oracle import cpu as O, ref as R
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    N, C, Ms = 5000, 24, 777
    perm = rng.permutation(N).astype(np.int32)
    gp, gf = rng.normal(size=(N, 3)).astype(np.float32), rng.normal(size=(N, C)).astype(np.float32)
    idx = np.sort(rng.choice(N, Ms, replace=False)).astype(np.int32)
    feats, g = rng.normal(size=(N, C)).astype(np.float32), rng.normal(size=(Ms, C)).astype(np.float32)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    ip, iF = M.sort_points_step2_grad(T(perm), T(gp), T(gf))
    wp, wF = O.sort_points_step2_grad(perm, gp, gf)
    assert np.array_equal(ip.cpu().numpy(), wp) and np.array_equal(iF.cpu().numpy(), wF)
    fg = M.get_sampled_features_grad(T(idx), T(feats), T(g))
    assert np.array_equal(fg.cpu().numpy(), O.get_sampled_features_grad(idx, feats, g))
    if R.available():
        rp, rF = R.sort_points_step2_grad(T(perm), T(gp), T(gf))
        rg = R.get_sampled_features_grad(T(idx), T(feats), T(g))
        R.lib().ref_sync()
        assert torch.equal(rp, ip) and torch.equal(rF, iF) and torch.equal(rg, fg)
    with pytest.raises(ValueError):
        M.get_sampled_features_grad(T(idx), T(feats), T(g[:, :3].copy()))


def test_autograd_through_builder_convolution():
    """loss.backward() through ConvolutionBuilder.create_convolution: feature re-sort -> spatial_conv, pooled samples
    from a Poisson hierarchy; gradients of the input features (in the caller's point order) and of the six parameter
    tensors (in their TF shapes) equal the oracle's chain SortFeaturesBack(SpatialConvGrad(...))."""
    import torch
    from mccnn_b200 import variables
    from mccnn_b200.MCConvBuilder import PointHierarchy, ConvolutionBuilder
    from oracle import cpu as O
    dev = torch.device("cuda:0")
    B, C, k, r = 2, 4, 6, 0.25
    pts, bids, feats = make_batch(B, 600, "sphere", seed=21, num_features=C)
    tp, tb = torch.from_numpy(pts).to(dev), torch.from_numpy(bids).to(dev)
    tf = torch.from_numpy(feats).to(dev).requires_grad_(True)
    store = variables.VariableStore(device=dev, seed=4)
    ph = PointHierarchy(tp, tf.detach(), tb, [0.1], "H", B)
    cb = ConvolutionBuilder(KDEWindow=0.25, variableStore=store)
    out = cb.create_convolution("c", ph, 0, tf, C, r, outPointLevel=1, multiFeatureConv=True, outNumFeatures=k)
    for v in store.vars.values():     # second pass with trainable, non-trivial parameters
        v.copy_(torch.from_numpy(np.random.default_rng(v.numel()).uniform(-0.4, 0.4, tuple(v.shape))
                                 .astype(np.float32)).to(dev))
        v.requires_grad_(True)
    cb.reset()
    out = cb.create_convolution("c", ph, 0, tf, C, r, outPointLevel=1, multiFeatureConv=True, outNumFeatures=k)
    G = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, tuple(out.shape)).astype(np.float32)).to(dev)
    (out * G).sum().backward()

    # oracle chain on the same geometry
    mn, mx = O.compute_aabb(pts, bids, B, True)
    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, r, True)
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, r, True)
    smp, smpb = ph.points_[1].cpu().numpy(), ph.batchIds_[1].cpu().numpy()
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, r, B, True)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.25, r, B, True)
    P = {n: v.detach().cpu().numpy() for n, v in store.vars.items()}
    nb = P["c_weights"].shape[1] // 8
    w = (P["c_weights"], P["c_biases"], P["c_weights2"].reshape(8, nb * 8), P["c_biases2"].reshape(-1),
         P["c_weights3"].reshape(8, nb * 8), P["c_biases3"].reshape(-1))
    args = (sp, sf, sb, pdf, smp, start, pairs, mn, mx, *w, G.cpu().numpy(), k, True, B, r, True, True)
    want, scale = O.spatial_conv_grad(*args), O.spatial_conv_grad(*args, abs_sum=True)
    assert _err(tf.grad.cpu().numpy(), O.sort_features_back(want[0], idx), O.sort_features_back(scale[0], idx)) <= 1e-5
    for name, i in [("c_weights", 1), ("c_biases", 2), ("c_weights2", 3), ("c_biases2", 4), ("c_weights3", 5),
                    ("c_biases3", 6)]:
        g = store.vars[name].grad
        assert g is not None and tuple(g.shape) == tuple(store.vars[name].shape), name
        assert _err(g.cpu().numpy().reshape(want[i].shape), want[i], scale[i]) <= 1e-5, name
