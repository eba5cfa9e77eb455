"""Full-size checks (BASELINE.json configs 2, 4 and 5: 32 x 1024, 16 x 8192, 8 rooms x 80 000 points) through
size-independent properties of the domain -- the oracle comparison of tests/test_parity_gpu.py is run at sizes a scalar
CPU port finishes in seconds; here the CUDA path is held to invariants that need no second implementation:

  grid sort      keys non-decreasing, new index a permutation, STABLE inside a cell, cell table = run boundaries
  neighbours     CSR consistent, every pair within r, none missed (exact brute force on sampled rows), symmetric
  Poisson disk   accepted points pairwise >= r apart, every rejected point within r of an accepted one, output ordered
  pdf            strictly positive
  spatial_conv   linear in the features, zero features -> zero, and <conv(f), G> = <f, conv_grad(G)> (adjointness ties
                 the backward kernel to the forward one at full size)
All comparisons of geometric predicates use the same float expression as the kernels (tests/common.ref_dist).
"""
import numpy as np
import pytest

from common import make_batch, make_conv_weights

pytestmark = pytest.mark.gpu

CONFIGS = [
    # name, batch, points per cloud, distribution, grid / Poisson radius, conv radius, scale_inv
    ("config2_mcclasss", 32, 1024, "sphere", 0.1, 0.2, True),
    ("config4_mcseg", 16, 8192, "sphere", 0.025, 0.05, True),
    ("config5_scannet", 8, 80000, "room", 0.1, 0.2, False),
]


def ref_dist(p, q):
    """sqrt(fma(dz,dz, fma(dx,dx, dy*dy))) in float32 -- the reference's compiled expression."""
    d = (q.astype(np.float32) - p.astype(np.float32)).astype(np.float32)
    dx, dy, dz = (d[..., k].astype(np.float64) for k in range(3))
    t = (dy * dy).astype(np.float32).astype(np.float64)            # FMUL, rounded
    t = (dx * dx + t).astype(np.float32).astype(np.float64)        # FFMA: exact product + add, rounded once
    t = (dz * dz + t).astype(np.float32)
    return np.sqrt(t, dtype=np.float32)


@pytest.fixture(scope="module", params=CONFIGS, ids=[c[0] for c in CONFIGS])
def scene(request):
    import torch
    import mccnn_b200.MCConvModule as M
    name, B, n, dist, r, rc, scaleInv = request.param
    dev = torch.device("cuda:0")
    pts, bids, feats = make_batch(B, n, dist, seed=17, num_features=8, ragged=(dist == "room"))
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    mn, mx = M.compute_aabb(tp, tb, B, scaleInv)
    keys, idx = M.sort_points_step1(tp, tb, mn, mx, B, r, scaleInv)
    sp, sb, sf, cells = M.sort_points_step2(tp, tb, tf, keys, idx, mn, mx, B, r, scaleInv)
    ext = (mx - mn).max(dim=1).values
    rad = (r * ext if scaleInv else torch.full_like(ext, r)).cpu().numpy()
    return dict(M=M, torch=torch, dev=dev, B=B, r=r, rc=rc, scaleInv=scaleInv, pts=pts, bids=bids, feats=feats,
                tp=tp, tb=tb, tf=tf, mn=mn, mx=mx, keys=keys, idx=idx, sp=sp, sb=sb, sf=sf, cells=cells, rad=rad,
                ext=ext.cpu().numpy())


def test_grid_sort_invariants(scene):
    s, torch = scene, scene["torch"]
    N = s["pts"].shape[0]
    keys, idx = s["keys"].cpu().numpy(), s["idx"].cpu().numpy()
    assert np.array_equal(np.sort(idx), np.arange(N))                      # a permutation
    skeys = np.empty_like(keys)
    skeys[idx] = keys
    assert (np.diff(skeys.astype(np.int64)) >= 0).all()                      # sorted by key
    order = np.argsort(idx)                                                  # original index at every sorted position
    same = skeys[1:] == skeys[:-1]
    assert (np.diff(order)[same] > 0).all()                                  # stable inside a cell
    assert np.array_equal(s["sp"].cpu().numpy(), s["pts"][order])
    assert np.array_equal(s["sb"].cpu().numpy(), s["bids"][order])
    assert np.array_equal(s["sf"].cpu().numpy(), s["feats"][order])
    cells = s["cells"].cpu().numpy().reshape(-1, 2)
    occ = cells[:, 1] > cells[:, 0]
    assert (cells[~occ] == 0).all()                                          # empty cells are (0, 0)
    assert int((cells[occ, 1] - cells[occ, 0]).sum()) == N
    uk, first = np.unique(skeys, return_index=True)                          # cell table = run boundaries of the keys
    assert np.array_equal(np.nonzero(occ)[0], uk)
    assert np.array_equal(cells[uk, 0], first)
    assert np.array_equal(cells[uk, 1], np.append(first[1:], N))
    # sort_features / sort_features_back are inverse permutations
    M = s["M"]
    assert torch.equal(M.sort_features_back(M.sort_features(s["tf"], s["idx"]), s["idx"]), s["tf"])


def test_poisson_disk_and_neighbour_invariants(scene):
    s, torch, M = scene, scene["torch"], scene["M"]
    B, r, scaleInv = s["B"], s["r"], s["scaleInv"]
    q, qb, qi = M.poisson_sampling(s["sp"], s["sb"], s["cells"], s["mn"], s["mx"], r, B, scaleInv)
    Ms, N = q.shape[0], s["sp"].shape[0]
    assert 0 < Ms < N
    qi_np, qb_np = qi.cpu().numpy(), qb.cpu().numpy().reshape(-1)
    assert len(np.unique(qi_np)) == Ms
    assert torch.equal(q, s["sp"][qi.long()]) and torch.equal(qb, s["sb"][qi.long()])
    assert (np.diff(qb_np) >= 0).all()                                       # output grouped by cloud, ascending
    # neighbours of every ORIGINAL point among the accepted ones (grid of the accepted set, same radius)
    k2, i2 = M.sort_points_step1(q, qb, s["mn"], s["mx"], B, r, scaleInv)
    sq, sqb, _, c2 = M.sort_points_step2(q, qb, q.new_zeros((Ms, 1)), k2, i2, s["mn"], s["mx"], B, r, scaleInv)
    start, pairs = M.find_neighbors(s["sp"], s["sb"], sq, c2, s["mn"], s["mx"], r, B, scaleInv)
    K = pairs.shape[0]
    st = start.cpu().numpy().reshape(-1)
    cnt = np.diff(np.append(st, K))
    assert st[0] == 0 and (cnt >= 0).all()
    assert (cnt >= 1).all()                       # maximality: every point is within r of an accepted point
    assert (cnt[qi_np] == 1).all()                # disk property: an accepted point sees only itself
    pj, pi = pairs[:, 0].cpu().numpy(), pairs[:, 1].cpu().numpy()
    assert (np.diff(pi) >= 0).all() and np.array_equal(np.repeat(np.arange(N), cnt), pi)   # CSR by sample
    spn, sqn, sbn = s["sp"].cpu().numpy(), sq.cpu().numpy(), s["sb"].cpu().numpy().reshape(-1)
    assert (ref_dist(spn[pi], sqn[pj]) < s["rad"][sbn[pi]]).all()            # every pair within r
    assert (sqb.cpu().numpy().reshape(-1)[pj] == sbn[pi]).all()               # never across clouds
    # completeness on sampled rows: exact brute force against all accepted points of the same cloud
    rng = np.random.default_rng(3)
    sqbn = sqb.cpu().numpy().reshape(-1)
    for i in rng.choice(N, 64, replace=False):
        cand = np.nonzero(sqbn == sbn[i])[0]
        inside = cand[ref_dist(spn[i][None, :], sqn[cand]) < s["rad"][sbn[i]]]
        assert np.array_equal(np.sort(pj[st[i]: st[i] + cnt[i]]), inside)


def test_neighbour_symmetry_pdf_and_conv_properties(scene):
    s, torch, M = scene, scene["torch"], scene["M"]
    B, rc, scaleInv, dev = s["B"], s["rc"], s["scaleInv"], s["dev"]
    # convolution grid at the conv radius; samples == points -> the neighbour relation is symmetric
    tp, tb = s["tp"], s["tb"]
    sub = slice(0, None, 1 if tp.shape[0] <= 200000 else 4)                  # rooms: every 4th point keeps K bounded
    tp, tb = tp[sub].contiguous(), tb[sub].contiguous()
    Nn = tp.shape[0]
    C = 8
    feats = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (Nn, C)).astype(np.float32)).to(dev)
    mn, mx = M.compute_aabb(tp, tb, B, scaleInv)
    keys, idx = M.sort_points_step1(tp, tb, mn, mx, B, rc, scaleInv)
    sp, sb, sf, cells = M.sort_points_step2(tp, tb, feats, keys, idx, mn, mx, B, rc, scaleInv)
    start, pairs = M.find_neighbors(sp, sb, sp, cells, mn, mx, rc, B, scaleInv)
    K = pairs.shape[0]
    assert K >= Nn                                                           # everybody neighbours itself
    code = pairs[:, 0].long() * Nn + pairs[:, 1].long()
    back = pairs[:, 1].long() * Nn + pairs[:, 0].long()
    assert torch.equal(torch.sort(code).values, torch.sort(back).values)     # (j, i) present <=> (i, j) present
    pdf = M.compute_pdf(sp, sb, mn, mx, start, pairs, 0.25, rc, B, scaleInv)
    assert pdf.shape == (K, 1) and bool((pdf > 0).all()) and bool(torch.isfinite(pdf).all())

    for combin, Co in ((False, C), (True, 4)):
        w = [torch.from_numpy(a).to(dev) for a in make_conv_weights(C, Co, combin, seed=9)]

        def conv(f):
            return M.spatial_conv(sp, f, sb, pdf, sp, start, pairs, mn, mx, *w, Co, combin, B, rc, scaleInv, True)

        f1 = sf
        f2 = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (Nn, C)).astype(np.float32)).to(dev)
        o1, o2, o12 = conv(f1), conv(f2), conv(0.5 * f1 - 2.0 * f2)
        scale = float(torch.maximum(o1.abs(), o2.abs()).max()) + 1e-30
        assert float((o12 - (0.5 * o1 - 2.0 * o2)).abs().max()) <= 2e-5 * scale      # linear in the features
        assert float(conv(torch.zeros_like(f1)).abs().max()) == 0.0
        # adjointness: sum(conv(f) * G) == sum(f * dF(G)), dF from the backward kernel
        G = torch.from_numpy(np.random.default_rng(4).uniform(-1, 1, tuple(o1.shape)).astype(np.float32)).to(dev)
        w_op = (w[0], w[3], w[1], w[4], w[2], w[5])                          # OP order W1, b1, W2, b2, Wout, bout
        dF = M.spatial_conv_grad(sp, f1, sb, pdf, sp, start, pairs, mn, mx, *w_op, G, Co, combin, B, rc, scaleInv,
                                 True)[0]
        lhs = float((o1.double() * G.double()).sum())
        rhs = float((f1.double() * dF.double()).sum())
        norm = float((o1.double().abs() * G.double().abs()).sum()) + 1e-30
        assert abs(lhs - rhs) <= 1e-5 * norm
