"""Per-op achieved bandwidth at a configuration where HBM matters (BASELINE.json config 5 by default: MCSegScanNet on
8 rooms x 80 000 points):  python tools/op_roofline.py [--model MCSegScanNet|MCSeg|MCClassH|MCClassS]

Every op of the reference-facing API (mccnn_b200.ops level) is bracketed by CUDA events during one planned (sync-free)
forward, its ALGORITHMIC bytes are evaluated from the live tensor sizes with SURVEY.md section 8(d)'s formulas, and the
table reports time, bytes, GB/s and the fraction of the measured copy bandwidth (MEASURED_PEAKS.json) per op kind."""
import argparse
import collections
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from common import make_batch
from mccnn_b200 import ops, runtime, variables
from mccnn_b200.engine import PlannedForward

CONFIGS = {
    "MCClassS": dict(B=32, N=1024, dist="sphere", k=64, args=lambda c: (1, c["B"], c["k"], 40, False, 1.0, 1.0)),
    "MCClassH": dict(B=32, N=4096, dist="sphere", k=64, args=lambda c: (1, c["B"], c["k"], 40, False, 1.0, 1.0)),
    "MCSeg": dict(B=16, N=8192, dist="sphere", k=64, args=lambda c: (1, 16, 50, c["B"], c["k"], False, 1.0, 1.0)),
    "MCSegScanNet": dict(B=8, N=80000, dist="room", k=32, args=lambda c: (1, 21, c["B"], c["k"], False, 1.0, 1.0)),
}


def live(t):
    c = runtime.count_of(t)
    return int(c.item()) if c is not None else int(t.shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="MCSegScanNet")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    c = CONFIGS[args.model]
    pts, bids, feats = make_batch(c["B"], c["N"], c["dist"], seed=3)
    t = [torch.from_numpy(a).to(dev) for a in (pts, bids, feats)]
    cat = torch.from_numpy((bids % 16).astype(np.int32)).to(dev)
    store = variables.VariableStore(device=dev, seed=0)
    mod = importlib.import_module("mccnn_b200.models." + args.model)

    def fn(p, b, f):
        lead = (p, b, f) + ((cat,) if args.model == "MCSeg" else ())
        return mod.create_network(*lead, *c["args"](c), store=store)

    rows = c["B"] if args.model.startswith("MCClass") else pts.shape[0]
    eng = PlannedForward(fn, rows, use_graph=False, multistream=False)
    eng(*t)
    eng.run_async(*t)
    torch.cuda.synchronize()

    records = []  # (op, start event, end event, bytes thunk)
    B = c["B"]

    def wrap(name, fnc, nbytes):
        def w(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fnc(*a, **kw)
            e.record()
            records.append((name, s, e, (nbytes, a, kw, r)))
            return r
        return w

    def G(cells):
        return int(cells.numel() // 2)

    byt = {
        "compute_aabb": lambda a, kw, r: 16 * a[0].shape[0] + 24 * B,
        "build_grid": lambda a, kw, r: live(a[0]) * (48 + 8 * a[2].shape[1]) + 20 * G(r[5]),
        "sort_points_step1": lambda a, kw, r: live(a[0]) * 24 + 12 * B * a[5] ** 3,
        "sort_points_step2": lambda a, kw, r: live(a[0]) * (24 + 8 * a[2].shape[1]) + 8 * G(r[3]),
        "sort_features": lambda a, kw, r: live(a[0]) * (8 * a[0].shape[1] + 4),
        "get_sampled_features": lambda a, kw, r: live(a[0]) * (4 + 8 * a[1].shape[1]),
        "find_neighbors_fused": lambda a, kw, r: (16 * live(a[0]) + 12 * live(a[2]) + 8 * min(G(a[3]), 27 * live(a[0]))
                                                 + 4 * live(a[0]) + 8 * int(r[2].item())),
        "compute_pdf": lambda a, kw, r: 12 * live(a[5]) + 4 * live(a[4]) + 12 * live(a[0]),
        "poisson_sampling": lambda a, kw, r: 13 * live(a[0]) + 8 * G(a[2]) + 20 * int(r[3].item()),
        "spatial_conv": lambda a, kw, r: (12 * live(a[6]) + 4 * a[1].shape[1] * live(a[6]) + 12 * live(a[0]) +
                                         16 * live(a[4]) + 4 * r.shape[1] * live(a[4])),
    }
    orig = {n: getattr(ops, n) for n in byt}
    for n in byt:
        setattr(ops, n, wrap(n, orig[n], byt[n]))
    torch.cuda._sleep(40_000_000)          # keep the GPU busy while the host enqueues the forward
    eng.run_async(*t)
    torch.cuda.synchronize()
    for n in byt:
        setattr(ops, n, orig[n])

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    agg = collections.OrderedDict()
    for name, s, e, (nb, a, kw, r) in records:
        us = s.elapsed_time(e) * 1e3
        d = agg.setdefault(name, [0, 0.0, 0])
        d[0] += 1
        d[1] += us
        d[2] += int(nb(a, kw, r))
    print("%s: %d points, peak %.0f GB/s" % (args.model, pts.shape[0], peak))
    print("%-22s %5s %10s %12s %9s %7s" % ("op", "calls", "us", "alg bytes", "GB/s", "frac"))
    for name, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = by / us * 1e-3
        print("%-22s %5d %10.1f %12d %9.1f %7.4f" % (name, n, us, by, gbs, gbs / peak))


if __name__ == "__main__":
    main()
