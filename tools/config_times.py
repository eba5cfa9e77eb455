"""Forward time of the other BASELINE.json configurations (parity-test cases, not bench lines) on one GPU:
    python tools/config_times.py [--configs MCClassH,MCSeg,MCSegScanNet] [--scale 1.0]
config 3: MCClassH 32 x 4096, k = 64;  config 4: MCSeg 16 x 8192, k = 64 (single GPU here);
config 5: MCSegScanNet 8 rooms x 80 000, k = 32.  Reports exact-mode (host syncs), planned and CUDA-graph replay times
and the per-op launch summary of one planned forward."""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from common import make_batch
from mccnn_b200 import ops, variables
from mccnn_b200.engine import PlannedForward

CONFIGS = {
    "MCClassS": dict(B=32, N=1024, dist="sphere", cin=1, k=64, args=lambda c: (1, c["B"], c["k"], 40, False, 1.0, 1.0)),
    "MCClassH": dict(B=32, N=4096, dist="sphere", cin=1, k=64, args=lambda c: (1, c["B"], c["k"], 40, False, 1.0, 1.0)),
    "MCSeg": dict(B=16, N=8192, dist="sphere", cin=1, k=64, args=lambda c: (1, 16, 50, c["B"], c["k"], False, 1.0, 1.0)),
    "MCSegScanNet": dict(B=8, N=80000, dist="room", cin=1, k=32, args=lambda c: (1, 21, c["B"], c["k"], False, 1.0, 1.0)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="MCClassH,MCSeg,MCSegScanNet")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the points per cloud")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for name in args.configs.split(","):
        try:
            run_config(name, args, dev)
        except Exception as e:  # keep going: one failing configuration must not hide the others
            import traceback
            traceback.print_exc()
            print({"config": name, "error": repr(e)[:300]}, flush=True)
            torch.cuda.synchronize()


def run_config(name, args, dev):
    if True:
        c = dict(CONFIGS[name])
        c["N"] = int(c["N"] * args.scale)
        pts, bids, feats = make_batch(c["B"], c["N"], c["dist"], seed=3, num_features=c["cin"])
        t = [torch.from_numpy(a).to(dev) for a in (pts, bids, feats)]
        cat = torch.from_numpy((bids % 16).astype(np.int32)).to(dev)
        store = variables.VariableStore(device=dev, seed=0)
        mod = importlib.import_module("mccnn_b200.models." + name)
        rows = c["B"] if name.startswith("MCClass") else pts.shape[0]

        def fn(p, b, f):
            lead = (p, b, f) + ((cat,) if name == "MCSeg" else ())
            return mod.create_network(*lead, *c["args"](c), store=store)

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*t)
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(3):
            out = fn(*t)
        torch.cuda.synchronize()
        exact = (time.perf_counter() - t0) / 3
        res = {"config": name, "points": int(pts.shape[0]), "first_s": round(first, 3), "exact_ms": round(exact * 1e3, 3)}
        for use_graph in (False, True):
            eng = PlannedForward(fn, rows, use_graph=use_graph)
            eng(*t)
            eng(*t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                o, flag = eng.run_async(*t)
            e1.record()
            torch.cuda.synchronize()
            res["graph_ms" if use_graph else "planned_ms"] = round(e0.elapsed_time(e1) / args.iters, 3)
            res["overflow"] = int(flag.item())
            if not use_graph:
                res["sizes"] = list(eng.plan.sizes)
        res["points_per_s_graph"] = round(pts.shape[0] / (res["graph_ms"] * 1e-3), 1)
        res["finite"] = bool(torch.isfinite(out).all())
        print(res, flush=True)
        # per-op wall of one exact forward: which op dominates
        names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and
                 n not in ("check", "new_count", "get_block_size", "num_cells", "build_grid_supported")]
        orig = {n: getattr(ops, n) for n in names}
        acc = {}

        def wrap(n):
            def w(*a, **kw):
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                r = orig[n](*a, **kw)
                torch.cuda.synchronize()
                d = acc.setdefault(n, [0, 0.0])
                d[0] += 1
                d[1] += time.perf_counter() - t1
                return r
            return w
        for n in names:
            if isinstance(orig[n], type(main)):
                setattr(ops, n, wrap(n))
        fn(*t)
        for n in names:
            setattr(ops, n, orig[n])
        tot = sum(v[1] for v in acc.values())
        print("   per-op (exact mode, synchronised): " + ", ".join(
            "%s x%d %.2f ms" % (n, v[0], v[1] * 1e3) for n, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:8]) +
            " | total %.2f ms" % (tot * 1e3), flush=True)


if __name__ == "__main__":
    main()
