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
