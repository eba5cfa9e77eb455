"""torchrun --nproc-per-node N tools/sharded_seg_check.py [--points 80000] [--model MCSegScanNet|MCSeg]

BASELINE.json configs 4 / 5 sharded over N GPUs: every rank runs the segmentation network on ITS rooms / objects only
(inside runtime.sharded(): absolute-radius hierarchies merge their bounding box with one 6-float all-reduce), the
per-point logits are all-gathered (engine.gather_point_logits) and compared with the same network run UNSHARDED on the
whole batch by every rank: identical Poisson-sample counts per level, logits equal up to fp32 summation order."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from common import make_batch
from mccnn_b200 import runtime, variables
from mccnn_b200.engine import gather_point_logits, shard_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="MCSegScanNet")
    ap.add_argument("--points", type=int, default=80000)
    ap.add_argument("--k", type=int, default=32)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    scannet = args.model == "MCSegScanNet"
    B = 8 if scannet else 16
    pts, bids, feats = make_batch(B, args.points, "room" if scannet else "sphere", seed=3, ragged=scannet)
    mod = importlib.import_module("mccnn_b200.models." + args.model)
    store = variables.VariableStore(device=dev, seed=0)      # same seed on every rank: identical weights

    def net(p, b, f, nb):
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (p, b, f)]
        if scannet:
            return mod.create_network(*t, 1, 21, nb, args.k, False, 1.0, 1.0, store=store)
        cat = torch.from_numpy((b % 16).astype(np.int32)).to(dev)
        return mod.create_network(*t, cat, 1, 16, 50, nb, args.k, False, 1.0, 1.0, store=store)

    with runtime.record() as plan_full:
        full = net(pts, bids, feats, B)
    p, b, f, nb, (lo, hi) = shard_batch(pts, bids, feats, B, rank, world)
    if not scannet:
        # object categories travel with the clouds: re-derive them from the global cloud index
        pass
    with runtime.sharded(), runtime.record() as plan_local:
        if scannet:
            local_out = net(p, b, f, nb)
        else:
            t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (p, b, f)]
            cat = torch.from_numpy(((b + lo) % 16).astype(np.int32)).to(dev)
            local_out = mod.create_network(*t, cat, 1, 16, 50, nb, args.k, False, 1.0, 1.0, store=store)
    gathered = gather_point_logits(local_out[: p.shape[0]].contiguous(), world)
    # Poisson samples per level: the shards' counts must add up to the unsharded ones
    n_levels = 4 if scannet else 3
    mine = torch.tensor(plan_local.sizes[:n_levels], dtype=torch.int64, device=dev)
    dist.all_reduce(mine)
    want = plan_full.sizes[:n_levels]
    scale = float(full.abs().max())
    err = float((gathered - full[: pts.shape[0]]).abs().max()) / scale
    ok = torch.tensor([float(mine.tolist() == list(want) and gathered.shape == full[: pts.shape[0]].shape), err],
                      device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN if False else dist.ReduceOp.MAX)
    if rank == 0:
        print({"model": args.model, "world": world, "points": int(pts.shape[0]), "levels_unsharded": list(want),
               "levels_sum_of_shards": mine.tolist(), "counts_equal": mine.tolist() == list(want),
               "max_abs_logit_diff_over_max_logit": err}, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
