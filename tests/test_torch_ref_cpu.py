"""tests/torch_ref.py (the float64 torch restatement of spatial_conv used by the full-size GPU parity tests) is held to
the C oracle -- itself pinned by the reference kernels' golden vectors -- on small cases: its value to the oracle's
float64-accumulated result, its error scale to the oracle's abs_sum mode."""
import numpy as np
import pytest
import torch

from common import make_batch, make_conv_weights
from torch_ref import spatial_conv_fp64

CASES = [
    # B, N, dist, C, Co, combin, radius, avg, scaleInv
    (1, 512, "sphere", 3, 8, True, 0.15, True, True),
    (2, 300, "sphere", 1, 16, True, 0.3, True, True),
    (2, 200, "cube", 16, 16, False, 0.5, False, True),
    (2, 300, "cube", 4, 5, True, 0.3, True, True),
    (2, 1500, "room", 4, 8, True, 0.4, True, False),
]


@pytest.mark.parametrize("B,N,dist,C,Co,combin,radius,avg,scaleInv", CASES)
def test_torch_fp64_conv_equals_the_oracle(B, N, dist, C, Co, combin, radius, avg, scaleInv):
    from oracle import cpu as O
    pts, bids, feats = make_batch(B, N, dist, seed=5, num_features=C)
    if C == 1:
        feats = np.random.default_rng(9).uniform(-1, 1, feats.shape).astype(np.float32)
    mn, mx = O.compute_aabb(pts, bids, B, scaleInv)
    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, radius, scaleInv)
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, radius, scaleInv)
    smp, smpb, _ = O.poisson_sampling(sp, sb, cells, mn, mx, radius * 0.5, B, scaleInv)
    start, pairs = O.find_neighbors(smp, smpb, sp, cells, mn, mx, radius, B, scaleInv)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.2, radius, B, scaleInv)
    w = make_conv_weights(C, Co, combin, seed=6)
    args = (sp, sf, sb, pdf, smp, start, pairs, mn, mx) + tuple(w)
    want = O.spatial_conv(*args, Co, combin, B, radius, scaleInv, avg, accum_double=True)
    want_scale = O.spatial_conv(*args, Co, combin, B, radius, scaleInv, avg, abs_sum=True)
    got, scale = spatial_conv_fp64(*[torch.from_numpy(np.ascontiguousarray(a)) for a in args], Co, combin, radius,
                                   scaleInv, avg, chunk=700)
    got, scale = got.numpy(), scale.numpy()
    assert got.shape == want.shape
    # the oracle evaluates the MLP in float32 and accumulates in float64; this is float64 throughout
    assert (np.abs(got - want) <= 2e-6 * want_scale + 1e-30).all()
    assert np.allclose(scale, want_scale, rtol=1e-5, atol=0)
