"""Development helper: per-op CUDA-event timing of one MCClassS forward (planned mode, no graph) and
whole-forward timing in exact / planned / graph mode, next to the reference kernels' op sequence.
Usage: python tools/quick_time.py [--batch 32] [--points 1024] [--k 64] [--iters 20]"""
import argparse
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from common import make_batch
import mccnn_b200
from mccnn_b200 import ops, variables
from mccnn_b200.engine import PlannedForward
from mccnn_b200.models import MCClassS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    store = variables.VariableStore(device=dev, seed=0)

    def fn(p, b, f):
        return MCClassS.create_network(p, b, f, 1, B, args.k, 40, False, 1.0, 1.0, store=store)

    pts, bids, feats = make_batch(B, args.points, "sphere", seed=0)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))

    def timeit(f, iters):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(iters):
            f()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters, (time.perf_counter() - t0) * 1e3 / iters

    print("exact   : %.3f ms gpu, %.3f ms wall" % timeit(lambda: fn(tp, tb, tf), args.iters))
    eng = PlannedForward(fn, B, use_graph=False)
    eng(tp, tb, tf)
    print("plan sizes", eng.plan.sizes, "caps", eng.plan.capacities)
    print("planned : %.3f ms gpu, %.3f ms wall" % timeit(lambda: eng.run_async(tp, tb, tf), args.iters))
    engg = PlannedForward(fn, B, use_graph=True)
    engg(tp, tb, tf)
    print("graph   : %.3f ms gpu, %.3f ms wall" % timeit(lambda: engg.run_async(tp, tb, tf), args.iters))

    # per-op timing in planned mode
    timings = collections.defaultdict(list)
    names = ["compute_aabb", "sort_points_step1", "sort_points_step2", "sort_features", "transform_indexs",
             "find_neighbors_count", "find_neighbors_pack", "compute_pdf", "poisson_sampling", "get_sampled_features",
             "spatial_conv"]
    orig = {n: getattr(ops, n) for n in names}
    events = []

    def wrap(n):
        def w(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig[n](*a, **kw)
            e.record()
            events.append((n, s, e))
            return r
        return w

    for n in names:
        setattr(ops, n, wrap(n))
    for it in range(args.iters):
        eng.run_async(tp, tb, tf)
    torch.cuda.synchronize()
    for n, s, e in events:
        timings[n].append(s.elapsed_time(e))
    for n in names:
        setattr(ops, n, orig[n])
    tot = 0.0
    for n in names:
        if timings[n]:
            per_fwd = sum(timings[n]) / args.iters
            tot += per_fwd
            print("  %-22s %8.1f us / forward (%d calls)" % (n, per_fwd * 1e3, len(timings[n]) // args.iters))
    print("  sum of ops %.1f us" % (tot * 1e3))

    # reference kernels, same op sequence
    from oracle import ref as R
    if R.available():
        import math
        def ref_forward():
            mn, mx = R.compute_aabb(tp, tb, B, True)
            levels = [(tp, tb, tf)]
            p, b, f = tp, tb, tf
            for r in [0.1, 0.4, math.sqrt(3.0) + 0.1]:
                keys, idx = R.sort_points_step1(p, b, mn, mx, B, r, True)
                sp, sb, sf, cells = R.sort_points_step2(p, b, f, keys, idx, mn, mx, B, r, True)
                qp, qb, qi = R.poisson_sampling(sp, sb, cells, mn, mx, r, B, True)
                qf = R.get_sampled_features(qi, sf)
                R.transform_indexs(qi, idx)
                levels.append((qp, qb, qf))
                p, b, f = qp, qb, qf
            return mn, mx, levels
        print("reference hierarchy only: %.3f ms gpu, %.3f ms wall" % timeit(ref_forward, max(3, args.iters // 4)))


if __name__ == "__main__":
    main()
