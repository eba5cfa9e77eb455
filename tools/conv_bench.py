"""Per-call timing of every spatial_conv (or dense layer) of a BASELINE configuration with the real inputs of a
forward:  python tools/conv_bench.py [--config 4] [--reps 20] [--once] [--op spatial_conv|dense_layer_tc]
The network runs once in exact mode while the arguments of every mccnn_b200.ops.<op> call are recorded; each call is
then replayed `reps` times between CUDA events (L2 flushed before every replay).  --once replays each call a single
time without timing (the target of an `ncu -k regex:spatial_conv_kernel` capture)."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
from mccnn_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--op", default="spatial_conv")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.config]
    dev = torch.device("cuda:0")
    pts, bids, feats = bench.host_batch(w, w["B"], seed=3)
    t = [torch.from_numpy(a).to(dev) for a in (pts, bids, feats)]
    lead = list(t)
    if w["model"] == "MCSeg":
        lead.append(torch.from_numpy((bids % w["cats"]).astype(np.int32)).to(dev))
    store = bench.named_store(dev)
    mod = importlib.import_module("mccnn_b200.models." + w["model"])
    calls = []
    orig = getattr(ops, args.op)

    def rec(*a, **kw):
        calls.append((a, kw))
        return orig(*a, **kw)
    setattr(ops, args.op, rec)
    if args.op in ("find_neighbors_fused", "find_neighbors_pdf_fused"):   # only issued by the planned (sync-free) path
        from mccnn_b200.engine import PlannedForward
        eng = PlannedForward(lambda p, b, f: mod.create_network(p, b, f, *lead[3:], *bench.net_args(w, w["B"], "ours"),
                                                                store=store),
                             w["B"] if bench.per_cloud_output(w) else pts.shape[0], use_graph=False, multistream=False)
        setattr(ops, args.op, orig)
        eng(*t)
        setattr(ops, args.op, rec)
        eng.run_async(*t)
    else:
        mod.create_network(*lead, *bench.net_args(w, w["B"], "ours"), store=store)
    setattr(ops, args.op, orig)
    torch.cuda.synchronize()
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peak = bench.measured_peaks()[0]
    total = 0.0
    total2 = 0.0
    print("%s %d x %d: %d %s calls" % (w["model"], w["B"], w["P"], len(calls), args.op))
    for n, (a, kw) in enumerate(calls):
        if args.once:
            orig(*a, **kw)
            continue
        for _ in range(2):
            orig(*a, **kw)
        us = []
        for _ in range(args.reps):
            flush.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(*a, **kw)
            e.record()
            torch.cuda.synchronize()
            us.append(s.elapsed_time(e) * 1e3)
        med = float(np.median(us))
        total += med
        if args.op == "spatial_conv":
            N, C, M, K = a[0].shape[0], a[1].shape[1], a[4].shape[0], a[6].shape[0]
            nout, combin = a[15], a[16]
            neurons = C * nout if combin else C
            nb = (neurons + 7) // 8
            by = 12 * K + 4 * C * K + 12 * N + 16 * M + 4 * nout * M
            print("  #%2d N=%7d M=%7d K=%8d (%5.1f / sample) C=%4d out=%4d blocks=%3d | %8.1f us | %7.1f GB/s alg (%.3f of "
                  "peak) | %6.2f G pair-blocks/s | %5.1f TFLOP/s useful" %
                  (n, N, M, K, K / max(M, 1), C, nout, nb, med, by / med * 1e-3, by / med * 1e-3 / peak,
                   K * nb / med * 1e-3, K * nb * 320 / med * 1e-6))
        elif args.op == "find_neighbors_pdf_fused":
            # the pair it replaces, on the same inputs: mc_find_neighbors_fused, then mc_compute_pdf on its result
            smp, sb, p2, cidx, mn, mx, rad, win, sinv, cap = a[:10]
            b2 = torch.zeros((p2.shape[0],), dtype=torch.int32, device=dev)
            # batch ids of the sorted points: cloud of the cell range they fall in is not needed for fixed radii;
            # with relative radii take them from the samples' neighbours
            out = orig(*a, **kw)
            k = int(out[3].item())
            if sinv:
                b2[out[1][:k, 0].long()] = sb[out[1][:k, 1].long()]

            def pair():
                st, pr, kd = ops.find_neighbors_fused(smp, sb, p2, cidx, mn, mx, rad, sinv, cap, *a[10:], **kw)
                return ops.compute_pdf(p2, b2, mn, mx, st, pr, win, rad, sinv, a[10] if len(a) > 10 else kw.get("m_dev"),
                                       kd)
            pair()
            us2 = []
            for _ in range(args.reps):
                flush.add_(1.0)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                pair()
                e.record()
                torch.cuda.synchronize()
                us2.append(s.elapsed_time(e) * 1e3)
            med2 = float(np.median(us2))
            total2 += med2
            print("  #%2d samples=%7d points=%7d pairs=%8d (%5.1f / sample) | fused %8.1f us | neighbours + pdf %8.1f us "
                  "| %.2fx" % (n, smp.shape[0], p2.shape[0], k, k / max(smp.shape[0], 1), med, med2, med2 / med))
        elif args.op.startswith("find_neighbors"):
            out = orig(*a, **kw)
            k = int(out[2].item()) if len(out) > 2 else -1
            print("  #%2d samples=%7d points=%7d cells/cloud=%6d pairs=%8d (%5.1f / sample) | %8.1f us | %6.1f M samples/s" %
                  (n, a[0].shape[0], a[2].shape[0], a[3].shape[1] ** 3, k, k / max(a[0].shape[0], 1), med,
                   a[0].shape[0] / med))
        elif args.op == "poisson_sampling":
            out = orig(*a, **kw)
            print("  #%2d points=%7d cells/cloud=%6d samples=%7d | %8.1f us | %6.1f M points/s" %
                  (n, a[0].shape[0], a[2].shape[1] ** 3, int(out[3].item()), med, a[0].shape[0] / med))
        elif args.op == "compute_pdf":
            print("  #%2d pairs=%8d samples=%7d | %8.1f us" % (n, a[5].shape[0], a[4].shape[0], med))
        else:
            x, wgt = a[0], a[1]
            tc = args.op.endswith("tc")
            print("  #%2d rows=%7d K=%5d N=%5d | %8.1f us | %6.1f TFLOP/s fp32-equivalent" %
                  (n, x.shape[0], wgt.shape[1] if tc else wgt.shape[0], wgt.shape[0] if tc else wgt.shape[1], med,
                   2.0 * x.shape[0] * wgt.shape[0] * wgt.shape[1] / med * 1e-6))
    if not args.once:
        print("  total %.1f us" % total)
        if total2 > 0:
            print("  total of the two separate launches %.1f us -> %.2fx" % (total2, total2 / total))


if __name__ == "__main__":
    main()
