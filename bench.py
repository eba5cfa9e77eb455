#!/usr/bin/env python
"""bench.py -- MCClassS forward points/s on synthetic 32 x 1024-point clouds (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one MCClassS forward (models/MCClassS.py topology: 3-level Poisson hierarchy, 3 Monte-Carlo convolutions,
dense glue, 40 logits per cloud) over one packed batch of 32 clouds x 1024 points per GPU.  Multi-GPU: independent
clouds sharded over the ranks (32 clouds per GPU, weak scaling), ONE NCCL all_gather of the per-cloud logits per step.

One JSON line on rank 0:
  value   points/s, whole job, batches already resident in HBM, CUDA-graph replay, CUDA-event timed per step
          (L2 flushed between steps, outside the timed intervals), max over ranks;
  e2e     the same metric through the host-facing call (pinned host buffers -> H2D -> forward -> D2H logits),
          wall clock around every call;
  roofline   dominant kernel of the step: algorithmic bytes / its mean CUDA-event duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
  cpu_baseline  the oracle port of the same forward on the host (1 thread) + NumPy spatial_conv of config 1;
`--impl reference` runs the reference's own CUDA kernels (oracle/_ref, built from /root/reference/tf_ops/*.cu) through
the same op sequence with the reference's host choreography (oracle/ref_pipeline.py) -- the reference has no CPU
implementation of this path; its GPU path re-compiled for sm_100a IS the reference arm.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

# Several forwards are in flight on (pipeline depth x 3) streams: ask the driver for the maximum number of hardware
# work queues before the CUDA context exists (default 8: streams that share a queue serialise -- measured 0.102 ms per
# step with 8 queues against 0.089 ms with 32 at depth 8)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from common import make_batch, make_conv_weights  # noqa: E402

NUM_CAT = 40
RADII = [0.1, 0.4, math.sqrt(3.0) + 0.1]
CONV_RADII = [0.2, 0.8, math.sqrt(3.0) + 0.1]


# ----------------------------------------------------------------------------------------------- utilities
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(op, call, sizes, B, k, n_in=1):
    """SURVEY.md section 8(d) per-op byte counts for call number `call` of `op` inside one MCClassS forward."""
    N = [sizes["N0"], sizes["M1"], sizes["M2"], sizes["M3"]]
    K = [sizes["K1"], sizes["K2"], sizes["K3"]]
    hier_n = [max(1, int(np.float32(1.0) / np.float32(r))) for r in RADII]
    conv_n = [max(1, int(np.float32(1.0) / np.float32(r))) for r in CONV_RADII]
    feat_c = [n_in, n_in, n_in]                      # features carried through the hierarchy
    conv_c = [n_in, 2 * k, 4 * k]
    conv_co = [k, 2 * k, 4 * k]
    if op == "compute_aabb":
        return 16 * N[0] + 24 * B
    if op == "sort_points":                          # calls: hier L0, hier L1, hier L2, conv1, conv2, conv3
        order = [("h", 0), ("h", 1), ("h", 2), ("c", 0), ("c", 1), ("c", 2)][call]
        lvl = order[1]
        n = hier_n[lvl] if order[0] == "h" else conv_n[lvl]
        return N[lvl] * 48 + 20 * B * n ** 3        # geometry only (C = 0): features move in sort_features
    if op == "sort_features":                        # calls: conv1, conv2, conv3
        return N[call] * (8 * conv_c[call] + 4)
    if op == "poisson_sampling":
        return 13 * N[call] + 8 * B * hier_n[call] ** 3 + 20 * N[call + 1]
    if op == "get_sampled_features":
        return N[call + 1] * (4 + 8 * feat_c[call])
    if op == "transform_indexs":
        return 8 * N[call] + 8 * N[call + 1]
    if op == "dense_layer":                          # Reduce_1, Reduce_2
        rows, kin, nout = [(N[1], k, 2 * k), (N[2], 2 * k, 4 * k)][call]
        return 4 * (rows * kin + kin * nout + rows * nout + 2 * kin + 3 * nout)
    if op == "mlp_rows":                             # the three layers of the head in one launch
        w = [4 * k, 2 * k, k, NUM_CAT]
        return 4 * (B * w[0] + B * w[3] + sum(a * b + 3 * b for a, b in zip(w[:-1], w[1:])) + 2 * w[0])
    if op in ("find_neighbors_count", "find_neighbors_pack", "find_neighbors", "find_neighbors_fused"):
        M, Np, G = N[call + 1], N[call], B * conv_n[call] ** 3
        full = 16 * M + 12 * Np + 8 * min(G, 27 * M) + 4 * M + 8 * K[call]
        return full if op != "find_neighbors_count" else full - 8 * K[call]
    if op == "compute_pdf":
        return 12 * K[call] + 4 * N[call + 1] + 12 * N[call]
    if op == "spatial_conv":
        C, Co = conv_c[call], conv_co[call]
        nOut = C * Co if call == 0 else C
        nb = (nOut + 7) // 8
        return (12 * K[call] + 4 * C * K[call] + 12 * N[call] + 12 * N[call + 1] + 4 * N[call + 1] +
                4 * Co * N[call + 1] + 4 * (8 * nb * 13 + nOut * 9))
    return 0


def flush_l2(buf):
    buf.add_(1.0)  # read + write a buffer larger than the 126 MB L2


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, dev):
    import torch.distributed as dist
    import mccnn_b200
    from mccnn_b200 import ops, variables
    from mccnn_b200._lib import lib
    from mccnn_b200.engine import (PlannedForward, PipelinedForward, HostForward, PipelinedHostForward,
                                   gather_logits, pack_batch)
    from mccnn_b200.models import MCClassS

    B, P, k = args.batch, args.points, args.k
    store = variables.VariableStore(device=dev, seed=0)

    def fn(p, b, f):
        return MCClassS.create_network(p, b, f, 1, B, k, NUM_CAT, False, 1.0, 1.0, store=store)

    n_batches = 4
    host = [make_batch(B, P, "sphere", seed=1000 * rank + i) for i in range(n_batches)]
    pinned = [tuple(torch.from_numpy(a).pin_memory() for a in h) for h in host]
    resident = [tuple(t.to(dev) for t in h) for h in pinned]

    eng = PlannedForward(fn, B, use_graph=True, slack=1.3)
    eng.calibrate(resident)
    sizes = dict(zip(["M1", "M2", "M3", "K1", "K2", "K3"], eng.plan.sizes))
    sizes["N0"] = B * P
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 192 MiB > L2

    def step(i):
        out, flag = eng.run_async(*resident[i % n_batches])
        if world > 1:
            out = gather_logits(out, B * world, world)
        return out, flag

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize(dev)

    # ---- value: device-resident inputs, per-step CUDA events, L2 flushed between steps
    sampler = ClockSampler(dev.index)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    flags = torch.zeros((1,), dtype=torch.int32, device=dev)
    launches0 = lib.mc_launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush_l2(flush)
        starts[i].record()
        out, flag = step(i)
        ends[i].record()
        flags += flag
    torch.cuda.synchronize(dev)
    t_wall = time.perf_counter() - t_wall0
    if world > 1:
        dist.barrier()
    per_step = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    ms_local = float(np.mean(per_step))
    overflowed = int(flags.item())
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    latency_ms = float(t.item())          # one forward at a time (L2 flushed before each)
    serial_value = B * P * world / (latency_ms * 1e-3)

    # ---- value: throughput with `depth` forwards in flight (independent CUDA graphs on `depth` streams).  The timed
    # region rotates over enough distinct resident batches that their inputs exceed the L2 (no flush possible inside
    # an overlapped region); one event pair around exactly K steps.
    depth = max(1, args.pipeline)
    n_pipe = max(n_batches, int(math.ceil(140e6 / (B * P * 20.0))))     # 20 B of input per point -> > 126 MB L2
    # resident batches in the single-block layout the captured graphs read (points | batch ids | features)
    pipe_res = [pack_batch(*r) for r in resident]
    for i in range(n_batches, n_pipe):
        pipe_res.append(pack_batch(*(torch.from_numpy(a).to(dev)
                                     for a in make_batch(B, P, "sphere", seed=1000 * rank + i))))
    # multi-GPU: the step's single collective (all_gather of the per-cloud logits) is captured INTO each lane's CUDA
    # graph (one NCCL communicator per lane: lanes run concurrently), so a step stays one graph launch on the host.
    # Fallback if the capture is refused: the collective is issued on the lane's stream right behind the replay.
    gather_mode, after = "none", None
    pipe = None
    if world > 1 and os.environ.get("MCCNN_BENCH_DIAG_NO_GATHER") == "1":
        gather_mode = "DIAGNOSTIC: collective skipped in the pipelined loops (not a valid bench line)"
        pipe = PipelinedForward(fn, B, depth=depth, slack=1.3)
        pipe.calibrate(resident)
    elif world > 1:
        # (1) own kernel over NVLink peer memory (csrc/peer_gather.cu), captured into each lane's graph;
        # (2) NCCL all_gather captured into each lane's graph (one communicator per lane); (3) NCCL per step
        want = os.environ.get("MCCNN_GATHER", "peer")
        if want == "peer":
            try:
                from mccnn_b200.engine import PeerGather
                gathers = [PeerGather(B, NUM_CAT, dev) for _ in range(depth)]
                pipe = PipelinedForward(fn, B, depth=depth, slack=1.3, epilogues=gathers)
                pipe.calibrate(resident)
                gather_mode = ("all_gather as ONE own kernel over NVLink peer memory (symmetric buffers, peer stores + "
                               "flags; csrc/peer_gather.cu), a node of each lane's CUDA graph")
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write("bench: peer-memory all_gather unavailable (%r); trying NCCL in the graph\n" % (exc,))
                pipe = None
        if pipe is None and want in ("peer", "nccl_graph"):
            try:
                groups = [dist.new_group(list(range(world))) for _ in range(depth)]
                pipe = PipelinedForward(fn, B, depth=depth, slack=1.3, epilogues=[
                    (lambda o, g=g: gather_logits(o, B * world, world, group=g)) for g in groups])
                pipe.calibrate(resident)
                gather_mode = "NCCL all_gather captured in each lane's CUDA graph (one communicator per lane)"
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write("bench: graph-captured all_gather unavailable (%r); issuing it per step\n" % (exc,))
                pipe = None
    if pipe is None:
        pipe = PipelinedForward(fn, B, depth=depth, slack=1.3)
        pipe.calibrate(resident)
        if world > 1 and not gather_mode.startswith("DIAG"):
            gather_mode = "all_gather issued on the lane's stream behind each graph replay"
            after = lambda out, flag: gather_logits(out, B * world, world)  # noqa: E731
    pflags = torch.zeros((1,), dtype=torch.int32, device=dev)

    def pstep(i):
        return pipe.submit_packed(pipe_res[i % n_pipe], after=after)

    for i in range(max(args.warmup, 3) * depth):
        pstep(i)
    pipe.join()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        pstep(i)
    pipe.join()
    ev1.record()
    torch.cuda.synchronize(dev)
    t_wall_pipe = time.perf_counter() - tw0
    if world > 1:
        dist.barrier()
    for e in pipe.engines:
        pflags += e.plan.overflow
    overflowed += int(pflags.item())
    t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    value = B * P * world / (ms_per_step * 1e-3)

    # ---- kernels per forward (ours) + per-op CUDA-event timing in eager planned mode (every op visible)
    eager = PlannedForward(fn, B, use_graph=False, slack=1.3, multistream=False)  # one stream: ops run back to back
    eager.plan, eager._shape_key = eng.plan, eng._shape_key
    eager.run_async(*resident[0])
    torch.cuda.synchronize(dev)
    c0 = lib.mc_launch_count()
    eager.run_async(*resident[0])
    launches_per_forward = int(lib.mc_launch_count() - c0)
    names = ["compute_aabb", "build_grid", "sort_points_step1", "sort_points_step2", "sort_features",
             "transform_indexs", "find_neighbors_count", "find_neighbors_pack", "find_neighbors_fused", "compute_pdf",
             "poisson_sampling", "get_sampled_features", "spatial_conv", "dense_layer", "mlp_rows", "scale_shift_act"]
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
    op_iters = min(args.steps, 10)
    for i in range(op_iters):
        flush_l2(flush)
        # keep the GPU busy (~6 ms) while the host enqueues the whole forward, so that every event pair brackets
        # kernel execution only and not the Python / allocator time between launches
        torch.cuda._sleep(12_000_000)
        eager.run_async(*resident[i % n_batches])
        torch.cuda.synchronize(dev)
    for n in names:
        setattr(ops, n, orig[n])
    per_call = {}
    counters = {}
    for idx, (n, s, e) in enumerate(events):
        c = counters.get((n, idx // (len(events) // op_iters)), 0)
        counters[(n, idx // (len(events) // op_iters))] = c + 1
        per_call.setdefault((n, c), []).append(s.elapsed_time(e) * 1e3)
    kernels = []
    for (n, c), v in per_call.items():
        opname = "sort_points" if n == "build_grid" else n
        by = algorithmic_bytes(opname, c, sizes, B, k)
        us = float(np.median(v))  # median over the timed forwards: robust against a one-off host hiccup
        kernels.append({"op": n, "call": c, "us": round(us, 2), "us_max": round(float(np.max(v)), 2),
                        "alg_bytes": int(by), "gbs": round(by / us * 1e-3, 2) if us > 0 else None})
    kernels.sort(key=lambda d: -d["us"])
    peak, peak_src = measured_peak_gbs()
    top = kernels[0]
    traffic = None
    prof = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                traffic = json.load(f).get(top["op"] + ":" + str(top["call"]))
        except Exception:
            traffic = None
    ncu = {}
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                ncu = json.load(f)
        except Exception:
            ncu = {}
    # spatial_conv against the tensor roofline (north_star: "tensor-pipe % of peak (spatial_conv)")
    conv = max((kk for kk in kernels if kk["op"] == "spatial_conv"), key=lambda kk: kk["us"], default=None)
    roofline_tensor = None
    if conv is not None:
        c = conv["call"]
        Kc = [sizes["K1"], sizes["K2"], sizes["K3"]][c]
        n_out = [k, 2 * k, 4 * k][c]                      # neurons: Cin*Co = k for Conv_1, C for the depth-wise ones
        nb = (n_out + 7) // 8
        flops = Kc * (176 * nb + 18 * n_out)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                tpeak = float(json.load(f)["bf16_tflops"])
            tsrc = "measured dense bf16 (MEASURED_PEAKS.json); the kernel computes fp32 via 3xTF32 mma.sync"
        except Exception:
            tpeak, tsrc = 2250.0, "nominal dense bf16"
        ach = flops / (conv["us"] * 1e-6) / 1e12
        roofline_tensor = {"bound": "tensor", "kernel": "spatial_conv (call %d of the forward)" % c,
                           "achieved": round(ach, 3), "peak": tpeak, "peak_source": tsrc, "unit": "TFLOP/s",
                           "frac": round(ach / tpeak, 5), "alg_flops_per_launch": int(flops),
                           "us_per_launch": conv["us"],
                           "tensor_pipe_active_pct_ncu": ncu.get("tensor_pipe_active_pct", {}).get(
                               "spatial_conv:%d" % c),
                           "note": "algorithmic FLOPs K*(176*nb + 18*Nout); 3xTF32 issues 3 MMAs per product and "
                                   "the Monte-Carlo sum adds 4 membership MMAs per tile, so the pipe-active % "
                                   "(ncu) is the utilisation figure, the TFLOP/s the useful-work figure"}
    roofline = {"bound": "hbm", "kernel": "%s (call %d of the forward)" % (top["op"], top["call"]),
                "achieved": top["gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(top["gbs"] / peak, 5) if top["gbs"] else None, "traffic": traffic,
                "alg_bytes_per_launch": top["alg_bytes"], "us_per_launch": top["us"],
                "note": "whole step is latency-bound: working set ~10 MB << L2; per-op table in 'kernels'"}

    # ---- e2e: host buffers, H2D + forward + D2H inside the timed call
    # (a) one call at a time (latency), (b) `depth` calls in flight (throughput) -- every step copies its inputs from
    # pinned host memory and its logits + overflow flag back, and every result is read on the host
    pinned_blocks = [pack_batch(*h, pin=True) for h in pinned]   # the data loader's pinned staging buffers
    hf = HostForward(eng, dev)
    for i in range(3):
        hf.call_packed(pinned_blocks[i % n_batches])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = hf.call_packed(pinned_blocks[i % n_batches])
    e2e_latency_s = (time.perf_counter() - t0) / args.steps
    phf = PipelinedHostForward(pipe, dev)
    checksum = 0.0
    for rep in range(2):
        warm = [phf.submit_packed(pinned_blocks[i % n_batches]) for i in range(depth)]
        for tk in warm:
            phf.result(tk)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    after_h = (lambda out: gather_logits(out, B * world, world)) if after is not None else None  # host gets ALL logits
    t0 = time.perf_counter()
    tickets = []
    for i in range(args.steps):
        tk = phf.submit_packed(pinned_blocks[i % n_batches], after=after_h)
        tickets.append(tk)
        if len(tickets) >= depth:                  # consume the oldest result before its lane is submitted to again
            res = phf.result(tickets.pop(0))
            checksum += float(res[0, 0])
    while tickets:
        res = phf.result(tickets.pop(0))
        checksum += float(res[0, 0])
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d, d2h = hf.bytes_per_call(*pinned[0], B, NUM_CAT)
    if world > 1:  # whole job: every rank copies its batch in and its result block + the gathered logits out
        h2d, d2h = h2d * world, (B * world * NUM_CAT * 4 + 4 + (d2h if after is not None else 0)) * world
    clocks = sampler.stop()  # sampled across the value, per-op and e2e regions

    line = None
    if rank == 0:
        cpu = cpu_baseline(store, host[0], B, k) if world == 1 else None  # reported at N = 1 only
        line = {
            "metric": "MC-conv fwd points/sec (MCClassS, 1024 pts, batch 32 per GPU)", "value": round(value, 1),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic sphere-surface clouds (seeded), random-init weights",
            "config": {"workload": "MCClassS forward (models/MCClassS.py), %d clouds x %d points per GPU, k=%d, "
                                   "%d classes, 1 input feature" % (B, P, k, NUM_CAT),
                       "parallelism": "dp%d (clouds sharded, one all_gather of logits)" % world,
                       "collective": gather_mode,
                       "concurrency": "%d forwards in flight per GPU (independent CUDA graphs on %d streams, "
                                      "CUDA_DEVICE_MAX_CONNECTIONS=%s); every step is a complete forward of its own "
                                      "batch" % (depth, depth, os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")),
                       "l2": "timed region rotates over %d distinct resident batches (%.0f MB of inputs > 126 MB L2); "
                             "the one-at-a-time latency run flushes L2 (192 MiB write) before every step"
                             % (n_pipe, n_pipe * B * P * 20.0 / 1e6),
                       "levels": [sizes["N0"], sizes["M1"], sizes["M2"], sizes["M3"]],
                       "pairs": [sizes["K1"], sizes["K2"], sizes["K3"]]},
            "clocks": clocks,
            "latency": {"ms_per_forward": round(latency_ms, 5), "points_per_s_one_at_a_time": round(serial_value, 1),
                        "note": "one forward at a time, L2 flushed before each, CUDA events per step"},
            "e2e": {"value": round(B * P * world / e2e_s, 1), "unit": "points/s", "ms_per_step": round(e2e_s * 1e3, 5),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "one_at_a_time": {"value": round(B * P * world / e2e_latency_s, 1),
                                      "ms_per_step": round(e2e_latency_s * 1e3, 5)},
                    "note": "%d host calls in flight; each copies its batch from pinned host memory, runs the forward "
                            "and returns logits + overflow flag to the host, every result is read on the host" % depth},
            "gpu_launches": (launches_per_forward + int("peer memory" in gather_mode)) * args.steps,
            "gpu_launches_per_step": launches_per_forward + int("peer memory" in gather_mode),
            "wall_s_timed_region": round(t_wall_pipe, 4),
            "plan_overflows": overflowed,
            "roofline": roofline,
            "roofline_tensor": roofline_tensor,
            "kernels": kernels[:12],
            "cpu_baseline": cpu,
        }
    return line


def cpu_baseline(store, host_batch, B, k):
    """Oracle port (oracle/mccnn_oracle.c + NumPy glue) of the same forward on the host, single thread, repeated for
    >= ~10 s; plus the NumPy re-execution of config 1's spatial_conv that BASELINE.json names."""
    from oracle import pipeline as OP, cpu as O, numpy_ops as NP
    params = {n: v.detach().cpu().numpy() for n, v in store.vars.items()}
    pts, bids, feats = host_batch
    t0 = time.perf_counter()
    reps = 0
    while True:
        OP.mcclasss_forward(pts, bids, feats, 1, B, k, NUM_CAT, params)
        reps += 1
        el = time.perf_counter() - t0
        if el > 10.0 or reps >= 20:
            break
    fwd_s = el / reps
    # config 1: single spatial_conv, 1 cloud x 1024 points, 3 -> 8 features, r = 0.1
    p1, b1, f1 = make_batch(1, 1024, "sphere", seed=0, num_features=3)
    mn, mx = O.compute_aabb(p1, b1, 1, True)
    keys, idx = O.sort_points_step1(p1, b1, mn, mx, 1, 0.1, True)
    sp, sb, sf, cells = O.sort_points_step2(p1, b1, f1, keys, idx, mn, mx, 1, 0.1, True)
    start, pairs = O.find_neighbors(sp, sb, sp, cells, mn, mx, 0.1, 1, True)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, 0.2, 0.1, 1, True)
    w = make_conv_weights(3, 8, True, seed=0)
    best = 1e9
    for _ in range(10):
        t1 = time.perf_counter()
        NP.spatial_conv(sp, sf, sb, pdf, sp, start, pairs, mn, mx, *w, 8, True, 1, 0.1, True, True)
        best = min(best, time.perf_counter() - t1)
    return {"value": round(pts.shape[0] / fwd_s, 1), "unit": "points/s", "cores": 1, "kind": "port",
            "sample": "full workload batch (%d clouds x %d points), %d repetitions, %.2f s per forward, "
                      "single-threaded C oracle + NumPy dense glue" % (B, pts.shape[0] // B, reps, fwd_s),
            "host_cores_available": os.cpu_count(),
            "numpy_spatial_conv_config1": {"ms": round(best * 1e3, 3), "pairs": int(pairs.shape[0]),
                                           "points_per_s": round(1024 / best, 1), "numpy_threads": "BLAS default",
                                           "note": "1 cloud x 1024 pts, 3->8 features, r=0.1, best of 10"}}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world, dev):
    """The reference's own CUDA kernels through the MCClassS op sequence (rank 0 only)."""
    if rank != 0:
        return None
    from oracle import ref as R
    if not R.available():
        return {"impl": "reference", "unavailable": "oracle/_ref/libmccnn_ref.so not built (needs /root/reference)"}
    from oracle import ref_pipeline as RP
    from mccnn_b200 import variables  # parameter initialisation only (same seed => same weights as our arm)
    from mccnn_b200.models import MCClassS
    B, P, k = args.batch, args.points, args.k
    store = variables.VariableStore(device=dev, seed=0)
    n_batches = 4
    host = [make_batch(B, P, "sphere", seed=i) for i in range(n_batches)]
    pinned = [tuple(torch.from_numpy(a).pin_memory() for a in h) for h in host]
    resident = [tuple(t.to(dev) for t in h) for h in pinned]
    # create the variables with our builder once (names/shapes == the reference's), then hand them over as a dict
    MCClassS.create_network(*resident[0], 1, B, k, NUM_CAT, False, 1.0, 1.0, store=store)
    params = dict(store.vars)

    def fwd(batch):
        return RP.mcclasss_forward(batch[0], batch[1], batch[2], 1, B, k, NUM_CAT, params)

    steps, warm = args.steps, max(args.warmup, 3)
    for i in range(warm):
        fwd(resident[i % n_batches])
    torch.cuda.synchronize(dev)
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms = []
    for i in range(steps):
        flush_l2(flush)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fwd(resident[i % n_batches])
        e.record()
        torch.cuda.synchronize(dev)
        ms.append(s.elapsed_time(e))
    clocks = sampler.stop()
    ms_per_step = float(np.mean(ms))
    # e2e: host buffers in, logits out
    dev_in = tuple(torch.empty_like(t, device=dev) for t in pinned[0])
    out_host = torch.empty((B, NUM_CAT), dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for i in range(steps):
        for d, h in zip(dev_in, pinned[i % n_batches]):
            d.copy_(h, non_blocking=True)
        out_host.copy_(fwd(dev_in)[:B])
        torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / steps
    value = B * P / (ms_per_step * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])
    return {"impl": "reference",
            "metric": "MC-conv fwd points/sec (MCClassS, 1024 pts, batch 32 per GPU)", "value": round(value, 1),
            "unit": "points/s", "n_gpus": 1, "launched_ranks": world, "steps": steps, "warmup": warm,
            "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic sphere-surface clouds (seeded), random-init weights",
            "config": {"workload": "MCClassS forward, %d clouds x %d points, k=%d, %d classes" % (B, P, k, NUM_CAT),
                       "path": "reference tf_ops CUDA kernels (unmodified, sm_100a) + TF-glue choreography shim"},
            "clocks": clocks,
            "e2e": {"value": round(B * P / e2e_s, 1), "unit": "points/s", "ms_per_step": round(e2e_s * 1e3, 4),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": B * NUM_CAT * 4},
            "cpu_baseline": {"value": round(value, 1), "unit": "points/s", "cores": 1, "kind": "reference",
                             "sample": "full workload batch per step; the reference has NO CPU implementation of this "
                                       "path -- this is its own CUDA path on the same GPU, driven by 1 host thread"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--pipeline", type=int, default=8, help="forwards in flight for the throughput numbers")
    args = ap.parse_args()

    # stdout carries exactly ONE line, the JSON: everything else any library prints there (e.g. NCCL's version banner)
    # is sent to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the step's collective moves ~5 KB per rank: one NCCL channel (one CTA) is plenty, and every extra channel is an
    # SM that spins while the lane waits for its peer
    os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("MCCNN_NCCL_CHANNELS", "1"))
    os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    if args.impl == "reference":
        line = run_reference(args, rank, world, dev)
        if line is not None:
            json_out.write(json.dumps(line) + "\n")
            json_out.flush()
        return

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    line = run_ours(args, rank, world, dev)
    if rank == 0 and line is not None:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
