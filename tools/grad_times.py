"""Backward op timing on the MCClassS convolution shapes: ours (mc_spatial_conv_grad) next to the reference's
spatialConvGradsCPU (oracle/_ref), CUDA events, same inputs.  python tools/grad_times.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from common import make_batch, make_conv_weights
import mccnn_b200.MCConvModule as M
from oracle import ref as R


def timeit(f, iters=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / iters


def main():
    dev = torch.device("cuda:0")
    B = 32
    pts, bids, _ = make_batch(B, 1024, "sphere", seed=0)
    tp, tb = torch.from_numpy(pts).to(dev), torch.from_numpy(bids).to(dev)
    mn, mx = M.compute_aabb(tp, tb, B, True)
    for name, C, Co, combin, r, stride in [("Conv_1 (1 -> 64, multi)", 1, 64, True, 0.2, 6),
                                           ("Conv_2-like (128, depth-wise)", 128, 128, False, 0.4, 12)]:
        feats = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (pts.shape[0], C)).astype(np.float32)).to(dev)
        keys, idx = M.sort_points_step1(tp, tb, mn, mx, B, r, True)
        sp, sb, sf, cells = M.sort_points_step2(tp, tb, feats, keys, idx, mn, mx, B, r, True)
        smp, smpb = sp[::stride].contiguous(), sb[::stride].contiguous()
        start, pairs = M.find_neighbors(smp, smpb, sp, cells, mn, mx, r, B, True)
        pdf = M.compute_pdf(sp, sb, mn, mx, start, pairs, 0.2, r, B, True)
        w = [torch.from_numpy(a).to(dev) for a in make_conv_weights(C, Co, combin, seed=2)]
        w_op = (w[0], w[3], w[1], w[4], w[2], w[5])
        outC = Co if combin else C
        G = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (smp.shape[0], outC)).astype(np.float32)).to(dev)
        geo = (sp, sf, sb, pdf, smp, start, pairs, mn, mx)
        attrs = (Co, combin, B, r, True, True)
        fwd = timeit(lambda: M.spatial_conv(*geo, *w, *attrs))
        ours = timeit(lambda: M.spatial_conv_grad(*geo, *w_op, G, *attrs))
        line = {"conv": name, "pairs": int(pairs.shape[0]), "samples": int(smp.shape[0]), "forward_us": round(fwd, 1),
                "backward_us": round(ours, 1)}
        if R.available():
            def ref():
                R.spatial_conv_grad(*geo, *w_op, G, *attrs)
            line["reference_backward_us"] = round(timeit(ref, 5), 1)
            line["speedup"] = round(line["reference_backward_us"] / ours, 1)
        print(line, flush=True)


if __name__ == "__main__":
    main()
