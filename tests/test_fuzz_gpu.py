"""Randomised parity sweep: 40 seeded configurations drawn over batch size, cloud size (ragged or not), point
distribution, radius, relative / absolute radius, feature counts and convolution kind -- the whole op chain of one
convolution (aabb -> sort -> Poisson samples -> neighbours -> pdf -> conv -> gradients) against the CPU oracle:
indices bit-exact, pdf 1e-5 relative, conv / gradients 1e-5 of their sums' |terms|."""
import numpy as np
import pytest

from common import make_batch, make_conv_weights, rel_err, sum_err

pytestmark = pytest.mark.gpu


def _draw(seed):
    rng = np.random.default_rng(1000 + seed)
    dist = ["cube", "sphere", "room", "lattice"][int(rng.integers(0, 4))]
    scaleInv = bool(rng.integers(0, 2)) if dist != "room" else False
    B = int(rng.integers(1, 7))
    N = int(rng.integers(40, 1500))
    if scaleInv:
        radius = float(rng.choice([0.05, 0.1, 0.17, 0.25, 0.4, 0.8, 1.9]))
    else:
        ext = 6.0 if dist == "room" else 2.0 if dist == "lattice" else 1.0
        radius = float(ext * rng.choice([0.05, 0.1, 0.125, 0.3]))
    combin = bool(rng.integers(0, 2))
    C = int(rng.choice([1, 2, 3, 5])) if combin else int(rng.choice([8, 16, 24]))
    Co = int(rng.integers(1, 12)) if combin else C
    while combin and (-(-(C * Co) // 8) * 8) % C != 0:   # op constraint: 8 * nb multiple of C (spatial_conv.cc:290)
        Co += 1
    return dict(dist=dist, scaleInv=scaleInv, B=B, N=N, radius=radius, combin=combin, C=C, Co=Co,
                ragged=bool(rng.integers(0, 2)), avg=bool(rng.integers(0, 2)),
                window=float(rng.choice([0.2, 0.25])), seed=seed)


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration(seed, monkeypatch):
    import torch
    # the sweep alternates over the kernels behind mc_spatial_conv (size-based choice, warp-level, persistent shapes)
    variant = ["auto", "warp", "8,2", "4,4", "4,2", "4,1"][seed % 6]
    if variant != "auto":
        monkeypatch.setenv("MCCNN_CONV_CFG", variant)
    else:
        monkeypatch.delenv("MCCNN_CONV_CFG", raising=False)
    import mccnn_b200.MCConvModule as M
    from oracle import cpu as O
    c = _draw(seed)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    A = lambda t: t.detach().cpu().numpy()  # noqa: E731
    B, r, si = c["B"], c["radius"], c["scaleInv"]
    pts, bids, feats = make_batch(B, c["N"], c["dist"], seed=50 + seed, num_features=c["C"], ragged=c["ragged"])
    w = make_conv_weights(c["C"], c["Co"], c["combin"], seed=seed)

    def chain(X, conv):
        mn, mx = X.compute_aabb(conv(pts), conv(bids), B, si)
        keys, idx = X.sort_points_step1(conv(pts), conv(bids), mn, mx, B, r, si)
        sp, sb, sf, cells = X.sort_points_step2(conv(pts), conv(bids), conv(feats), keys, idx, mn, mx, B, r, si)
        smp, smpb, sidx = X.poisson_sampling(sp, sb, cells, mn, mx, r * 0.5, B, si)
        start, pairs = X.find_neighbors(smp, smpb, sp, cells, mn, mx, r, B, si)
        pdf = X.compute_pdf(sp, sb, mn, mx, start, pairs, c["window"], r, B, si)
        out = X.spatial_conv(sp, sf, sb, pdf, smp, start, pairs, mn, mx, *[conv(a) for a in w], c["Co"], c["combin"], B,
                             r, si, c["avg"])
        return dict(mn=mn, mx=mx, keys=keys, idx=idx, sp=sp, sb=sb, sf=sf, cells=cells, smp=smp, smpb=smpb, sidx=sidx,
                    start=start, pairs=pairs, pdf=pdf, out=out)

    g = chain(M, T)
    o = chain(O, lambda a: a)
    for name in ("mn", "mx", "keys", "idx", "sp", "sb", "sf", "cells", "smp", "smpb", "sidx", "start", "pairs"):
        assert np.array_equal(A(g[name]).reshape(-1), np.asarray(o[name]).reshape(-1)), (name, c)
    if o["pairs"].shape[0] == 0:
        return
    assert rel_err(A(g["pdf"]), o["pdf"], 0.0) <= 1e-5, c
    args = (o["sp"], o["sf"], o["sb"], o["pdf"], o["smp"], o["start"], o["pairs"], o["mn"], o["mx"])
    scale = O.spatial_conv(*args, *w, c["Co"], c["combin"], B, r, si, c["avg"], abs_sum=True)
    assert sum_err(A(g["out"]), o["out"], scale) <= 1e-5, c
    # gradients on the same geometry
    G = np.random.default_rng(seed).uniform(-1, 1, o["out"].shape).astype(np.float32)
    w_op = (w[0], w[3], w[1], w[4], w[2], w[5])
    attrs = (c["Co"], c["combin"], B, r, si, c["avg"])
    want = O.spatial_conv_grad(*args, *w_op, G, *attrs)
    gscale = O.spatial_conv_grad(*args, *w_op, G, *attrs, abs_sum=True)
    got = M.spatial_conv_grad(*[T(a) for a in args], *[T(a) for a in w_op], T(G), *attrs)
    for k, (a, b, s) in enumerate(zip(got, want, gscale)):
        err = np.abs(A(a).astype(np.float64) - b) / (np.maximum(s, np.abs(b).mean()) + 1e-30)
        assert err.max() <= 1e-5, (k, c)


def _adversarial(kind, rng):
    """Small clouds built to hit the corners: heavy duplication, denormal-scale jitter, flat and collinear sets,
    one-point clouds, an absent cloud id in the middle of the batch."""
    if kind == "clumps":          # 6 sites x many exact duplicates (one cell holds > 32 points: chunked paths)
        sites = rng.uniform(-1, 1, (6, 3)).astype(np.float32)
        pts = sites[rng.integers(0, 6, 400)]
        bids = np.sort(rng.integers(0, 2, 400)).astype(np.int32)
        return pts, bids, 2
    if kind == "jitter":          # duplicates separated by ~1 ulp: distance 0 vs tiny
        base = rng.uniform(-1, 1, (40, 3)).astype(np.float32)
        pts = np.repeat(base, 8, axis=0)
        pts = np.nextafter(pts, np.float32(2.0) * rng.integers(0, 2, pts.shape).astype(np.float32) - 1).astype(np.float32)
        return pts, np.zeros(pts.shape[0], np.int32), 1
    if kind == "plane":
        pts = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
        pts[:, 2] = 0.25
        return pts, np.sort(rng.integers(0, 3, 500)).astype(np.int32), 3
    if kind == "line":
        t = rng.uniform(-1, 1, (300, 1)).astype(np.float32)
        pts = (t * np.array([[1.0, 0.5, -0.25]], np.float32)).astype(np.float32)
        return pts, np.zeros(300, np.int32), 1
    if kind == "singletons":      # clouds of 1-3 points, cloud id 2 absent (aabb rows stay +-FLT_MAX)
        pts = rng.uniform(-1, 1, (9, 3)).astype(np.float32)
        bids = np.array([0, 1, 1, 3, 3, 3, 4, 5, 5], np.int32)
        return pts, bids, 6
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["clumps", "jitter", "plane", "line", "singletons"])
@pytest.mark.parametrize("scaleInv,radius", [(True, 0.25), (False, 0.3)])
def test_adversarial_geometry_bit_exact(kind, scaleInv, radius):
    import torch
    import mccnn_b200.MCConvModule as M
    from oracle import cpu as O
    rng = np.random.default_rng(hash(kind) % 1000)
    pts, bids, B = _adversarial(kind, np.random.default_rng(sum(map(ord, kind))))
    bids = bids.reshape(-1, 1)
    feats = rng.uniform(-1, 1, (pts.shape[0], 2)).astype(np.float32)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    A = lambda t: t.detach().cpu().numpy()  # noqa: E731

    def chain(X, conv):
        mn, mx = X.compute_aabb(conv(pts), conv(bids), B, scaleInv)
        keys, idx = X.sort_points_step1(conv(pts), conv(bids), mn, mx, B, radius, scaleInv)
        sp, sb, sf, cells = X.sort_points_step2(conv(pts), conv(bids), conv(feats), keys, idx, mn, mx, B, radius,
                                                scaleInv)
        smp, smpb, sidx = X.poisson_sampling(sp, sb, cells, mn, mx, radius, B, scaleInv)
        start, pairs = X.find_neighbors(smp, smpb, sp, cells, mn, mx, radius, B, scaleInv)
        start2, pairs2 = X.find_neighbors(sp, sb, sp, cells, mn, mx, radius, B, scaleInv)
        return [mn, mx, keys, idx, sp, sb, sf, cells, smp, smpb, sidx, start, pairs, start2, pairs2]

    if kind == "singletons" and scaleInv:
        # one-point clouds have a zero-extent box (cell size 0): the cell of a point is 0/0 there -- defined by the
        # GPU's float-to-int conversion, not by C; only the absolute-radius variant is comparable on the CPU
        pytest.skip("zero-extent boxes: NaN cell coordinates, undefined in the C oracle")
    g = chain(M, T)
    o = chain(O, lambda a: a)
    names = ["mn", "mx", "keys", "idx", "sp", "sb", "sf", "cells", "smp", "smpb", "sidx", "start", "pairs", "start2",
             "pairs2"]
    for name, a, b in zip(names, g, o):
        assert np.array_equal(A(a).reshape(-1), np.asarray(b).reshape(-1)), (kind, name)
