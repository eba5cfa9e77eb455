"""Profiling target: MCClassS forward in planned mode (eager launches, every kernel visible to ncu).
Run under:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
                --log-file gpurun_out/launches.csv python tools/ncu_target.py --forwards 2
The profiled region (cudaProfilerStart/Stop) covers `--forwards` forwards after warm-up."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

from common import make_batch
from mccnn_b200 import variables
from mccnn_b200.engine import PlannedForward
from mccnn_b200.models import MCClassS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--forwards", type=int, default=2)
    ap.add_argument("--dist", default="sphere")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    store = variables.VariableStore(device=dev, seed=0)

    def fn(p, b, f):
        return MCClassS.create_network(p, b, f, 1, B, args.k, 40, False, 1.0, 1.0, store=store)

    pts, bids, feats = make_batch(B, args.points, args.dist, seed=0)
    tp, tb, tf = (torch.from_numpy(a).to(dev) for a in (pts, bids, feats))
    eng = PlannedForward(fn, B, use_graph=False)
    eng(tp, tb, tf)
    for _ in range(3):
        eng.run_async(tp, tb, tf)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.forwards):
        eng.run_async(tp, tb, tf)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("plan sizes", eng.plan.sizes)


if __name__ == "__main__":
    main()
