"""Recorded, sync-free, CUDA-graph-captured forward passes and their data-parallel sharding.

`PlannedForward` wraps any network function written against the reference API
(fn(points, batchIds, features) -> tensor, e.g. models.MCClassS.create_network):

  call 1      runs it in exact mode (host syncs like the reference) while recording every data-dependent size;
  later calls run the SAME Python code in planned mode -- capacity-sized tensors, device-resident counts, no host
              sync -- captured once into a CUDA graph and replayed: one graph launch per forward instead of
              ~2 700 kernel launches + >= 6 blocking syncs in the reference.
If an input ever needs more rows than planned, a device flag comes back with the result; the batch is then
re-run exactly and the plan is re-recorded (results are never silently truncated).

`shard_batch` / `gather_logits` implement the multi-GPU partition of BASELINE.json's north_star: independent clouds
of a packed batch are split contiguously over the ranks and ONE all_gather of the per-cloud logits closes the step.
"""
import torch

from . import runtime as _rt


class PlannedForward:
    def __init__(self, fn, out_rows, use_graph=True, slack=1.3, multistream=True):
        """fn(points, batchIds, features) -> [>= out_rows, F] tensor; out_rows rows of it are the result.
        multistream: overlap the hierarchy / convolution-geometry / feature chains on three streams (runtime.planned)."""
        self.fn = fn
        self.multistream = multistream
        self.out_rows = out_rows
        self.use_graph = use_graph
        self.slack = slack
        self.plan = None
        self.graph = None
        self._static_in = None
        self._static_out = None
        self._shape_key = None
        self.stats = {"records": 0, "captures": 0, "replays": 0, "overflows": 0}

    # -- exact, recording run ----------------------------------------------------------------------------------
    def _record(self, pts, bids, feats):
        with _rt.record(slack=self.slack) as plan:
            out = self.fn(pts, bids, feats)
        self.plan = plan
        self.graph = None
        self._shape_key = (tuple(pts.shape), tuple(feats.shape))
        self.stats["records"] += 1
        return out[: self.out_rows]

    def _planned(self, pts, bids, feats):
        with _rt.planned(self.plan, pts.device, multistream=self.multistream):
            return self.fn(pts, bids, feats)

    def _capture(self, pts, bids, feats):
        self._static_in = (pts.clone(), bids.clone(), feats.clone())
        self.plan.begin(pts.device)
        side = torch.cuda.Stream(device=pts.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture (allocator, lazy module loading)
            self._planned(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_out = self._planned(*self._static_in)
        self.graph = g
        self.stats["captures"] += 1

    # -- public ------------------------------------------------------------------------------------------------
    def run_async(self, pts, bids, feats):
        """Enqueue one forward; returns (out[out_rows, F], overflow_flag int32[1]) device tensors, no host sync.
        The first call (and any call with new input shapes) records and therefore does synchronise."""
        key = (tuple(pts.shape), tuple(feats.shape))
        if self.plan is None or key != self._shape_key:
            out = self._record(pts, bids, feats)
            return out, torch.zeros((1,), dtype=torch.int32, device=pts.device)
        if self.plan.overflow is None:
            self.plan.begin(pts.device)
        self.plan.overflow.zero_()
        if not self.use_graph:
            out = self._planned(pts, bids, feats)
            return out[: self.out_rows], self.plan.overflow
        if self.graph is None:
            self._capture(pts, bids, feats)
        sp, sb, sf = self._static_in
        sp.copy_(pts, non_blocking=True)
        sb.copy_(bids, non_blocking=True)
        sf.copy_(feats, non_blocking=True)
        self.graph.replay()
        self.stats["replays"] += 1
        return self._static_out[: self.out_rows], self.plan.overflow

    def __call__(self, pts, bids, feats):
        """Forward with the overflow check (one D2H of the flag); exact re-run + re-plan when a capacity was hit."""
        out, flag = self.run_async(pts, bids, feats)
        if int(flag.item()) != 0:
            self.stats["overflows"] += 1
            out = self._record(pts, bids, feats)
        return out

    def calibrate(self, batches):
        """Record several representative batches (same shapes) and plan for the element-wise maximum of their sizes."""
        sizes, flags = None, None
        for pts, bids, feats in batches:
            self._record(pts, bids, feats)
            s, f = self.plan.sizes, self.plan.flags
            sizes = list(s) if sizes is None else [max(a, b) for a, b in zip(sizes, s)]
            flags = list(f) if flags is None else [a and b for a, b in zip(flags, f)]
        self.plan.sizes, self.plan.flags = sizes, flags
        self.plan.finalize()
        self.graph = None
        return self.plan


class HostForward:
    """End-to-end call for host-resident batches: pinned host buffers -> H2D -> planned forward -> D2H of the result
    (and of the overflow flag) -> host tensor.  This is the user-facing equivalent of the reference's
    `sess.run(logits, feed_dict)` (ModelNet/ModelNetEval.py:123-125)."""

    def __init__(self, engine, device):
        self.engine = engine
        self.device = torch.device(device)
        self._dev = None
        self._out_host = None
        self._flag_host = torch.zeros((1,), dtype=torch.int32).pin_memory()

    def __call__(self, pts_host, bids_host, feats_host):
        """Arguments: CPU tensors (ideally pinned).  Returns a pinned CPU tensor [out_rows, F] (reused across calls)."""
        if self._dev is None or tuple(self._dev[0].shape) != tuple(pts_host.shape) or \
                tuple(self._dev[2].shape) != tuple(feats_host.shape):
            self._dev = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device)
                              for t in (pts_host, bids_host, feats_host))
        for d, h in zip(self._dev, (pts_host, bids_host, feats_host)):
            d.copy_(h, non_blocking=True)
        out, flag = self.engine.run_async(*self._dev)
        if self._out_host is None or tuple(self._out_host.shape) != tuple(out.shape):
            self._out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        self._out_host.copy_(out, non_blocking=True)
        self._flag_host.copy_(flag, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if int(self._flag_host[0]) != 0:  # a planned capacity was exceeded: exact re-run, re-plan
            self.engine.stats["overflows"] += 1
            self._out_host.copy_(self.engine._record(*self._dev))
        return self._out_host

    def bytes_per_call(self, pts_host, bids_host, feats_host, out_rows, out_cols):
        h2d = sum(t.numel() * t.element_size() for t in (pts_host, bids_host, feats_host))
        d2h = out_rows * out_cols * 4 + 4
        return h2d, d2h


class PipelinedForward:
    """Throughput mode: `depth` independent forwards in flight on one GPU.

    One MCClassS-sized forward is latency bound and occupies a fraction of the 148 SMs at any moment (a chain of small
    kernels); a server that has several batches queued can overlap them.  This class keeps `depth` PlannedForward
    engines -- each with its own CUDA graph and static buffers, all sharing the recorded plan sizes and the weights --
    and replays them round-robin on `depth` streams.  Every forward does the full work of the sequential one and
    produces bit-identical results; only the scheduling changes.

    submit(pts, bids, feats) enqueues one forward on the next lane and returns its ticket; result(ticket) waits for
    that forward only and returns (out, overflow_flag) device tensors (valid until the lane is reused)."""

    def __init__(self, fn, out_rows, depth=4, slack=1.3, multistream=True):
        self.depth = depth
        self.engines = [PlannedForward(fn, out_rows, use_graph=True, slack=slack, multistream=multistream)
                        for _ in range(depth)]
        self.streams = None
        self._events = [None] * depth
        self._results = [None] * depth
        self._next = 0

    def calibrate(self, batches):
        plan = self.engines[0].calibrate(batches)
        for e in self.engines[1:]:  # same sizes, own Plan object (own overflow flag, own graph)
            p = _rt.Plan(slack=plan.slack, granule=plan.granule)
            p.sizes, p.flags = list(plan.sizes), list(plan.flags)
            p.finalize()
            e.plan, e._shape_key, e.graph = p, self.engines[0]._shape_key, None
        dev = batches[0][0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
        for e, s in zip(self.engines, self.streams):  # capture every lane's graph up front
            with torch.cuda.stream(s):
                e.run_async(*batches[0])
        torch.cuda.synchronize(dev)
        return plan

    def submit(self, pts, bids, feats):
        lane = self._next
        self._next = (self._next + 1) % self.depth
        s = self.streams[lane]
        s.wait_stream(torch.cuda.current_stream())  # inputs produced on the caller's stream
        with torch.cuda.stream(s):
            out, flag = self.engines[lane].run_async(pts, bids, feats)
            ev = torch.cuda.Event()
            ev.record(s)
        self._events[lane], self._results[lane] = ev, (out, flag)
        return lane

    def result(self, ticket, wait_on_host=False):
        ev = self._events[ticket]
        if wait_on_host:
            ev.synchronize()
        else:
            torch.cuda.current_stream().wait_event(ev)
        return self._results[ticket]

    def join(self):
        for ev in self._events:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)


class PipelinedHostForward:
    """End-to-end throughput mode for host-resident batches: each lane of a PipelinedForward owns pinned staging, so
    the H2D copy of batch i+1, the forward of batch i and the D2H of batch i-1 overlap.  submit() takes CPU tensors
    (ideally pinned) and returns a ticket; result(ticket) blocks until that batch's logits are on the host."""

    def __init__(self, pipe, device):
        self.pipe = pipe
        self.device = torch.device(device)
        d = pipe.depth
        self._dev_in = [None] * d
        self._out_host = [None] * d
        self._flag_host = [torch.zeros((1,), dtype=torch.int32).pin_memory() for _ in range(d)]
        self._events = [None] * d
        self._inputs = [None] * d

    def submit(self, pts_host, bids_host, feats_host):
        lane = self.pipe._next
        if self._events[lane] is not None:
            self._events[lane].synchronize()  # the lane's previous result has reached the host
        s = self.pipe.streams[lane]
        if self._dev_in[lane] is None or tuple(self._dev_in[lane][0].shape) != tuple(pts_host.shape):
            self._dev_in[lane] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device)
                                       for t in (pts_host, bids_host, feats_host))
        with torch.cuda.stream(s):
            for dst, src in zip(self._dev_in[lane], (pts_host, bids_host, feats_host)):
                dst.copy_(src, non_blocking=True)
        ticket = self.pipe.submit(*self._dev_in[lane])
        out, flag = self.pipe._results[ticket]
        if self._out_host[lane] is None or tuple(self._out_host[lane].shape) != tuple(out.shape):
            self._out_host[lane] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        with torch.cuda.stream(s):
            self._out_host[lane].copy_(out, non_blocking=True)
            self._flag_host[lane].copy_(flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        self._events[lane] = ev
        self._inputs[lane] = (pts_host, bids_host, feats_host)
        return ticket

    def result(self, ticket):
        self._events[ticket].synchronize()
        if int(self._flag_host[ticket][0]) != 0:  # planned capacity exceeded: exact re-run of this batch
            eng = self.pipe.engines[ticket]
            eng.stats["overflows"] += 1
            with torch.cuda.stream(self.pipe.streams[ticket]):
                self._out_host[ticket].copy_(eng._record(*self._dev_in[ticket]))
            self.pipe.streams[ticket].synchronize()
        return self._out_host[ticket]


# ---- data-parallel sharding (one process per GPU) ---------------------------------------------------------------

def shard_range(batch_size, rank, world_size):
    """Contiguous cloud range [lo, hi) of `rank`; the first (batch_size % world_size) ranks get one extra cloud."""
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(points, batch_ids, features, batch_size, rank, world_size):
    """Slice a packed, cloud-contiguous batch (utils/DataSet.py:816-818) down to the clouds of `rank` and re-base the
    batch ids to 0..local_batch-1.  Works on numpy arrays and torch tensors alike.  Returns
    (points, batch_ids, features, local_batch_size, (lo, hi))."""
    lo, hi = shard_range(batch_size, rank, world_size)
    flat = batch_ids.reshape(-1)
    mask = (flat >= lo) & (flat < hi)
    return points[mask], (batch_ids[mask] - lo), features[mask], hi - lo, (lo, hi)


def gather_logits(local_logits, batch_size, world_size, group=None):
    """The single collective of the step: all_gather of per-cloud logits [local_batch, F] -> [batch_size, F]
    (rows padded to the largest shard so that every rank contributes the same size)."""
    import torch.distributed as dist
    if world_size == 1:
        return local_logits
    rows = (batch_size + world_size - 1) // world_size
    padded = local_logits.new_zeros((rows, local_logits.shape[1]))
    padded[: local_logits.shape[0]] = local_logits
    out = local_logits.new_empty((world_size * rows, local_logits.shape[1]))
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world_size):
        lo, hi = shard_range(batch_size, r, world_size)
        pieces.append(out[r * rows: r * rows + (hi - lo)])
    return torch.cat(pieces, dim=0)
