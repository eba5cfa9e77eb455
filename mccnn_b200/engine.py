"""Recorded, sync-free, CUDA-graph-captured forward passes, their pipelining and their data-parallel sharding.

`PlannedForward` wraps any network function written against the reference API
(fn(points, batchIds, features) -> tensor, e.g. models.MCClassS.create_network):

  call 1      runs it in exact mode (host syncs like the reference) while recording every data-dependent size;
  later calls run the SAME Python code in planned mode -- capacity-sized tensors, device-resident counts, no host
              sync -- captured once into a CUDA graph and replayed: one graph launch per forward instead of
              ~2 700 kernel launches + >= 6 blocking syncs in the reference.
If an input ever needs more rows than planned, a device flag comes back with the result; the batch is then
re-run exactly and the plan is re-recorded (results are never silently truncated).

The graph reads its three inputs from ONE contiguous static block and leaves logits + overflow flag in ONE contiguous
result block, so a host-resident batch costs exactly one H2D copy, one graph launch and one D2H copy
(`HostForward`).  `PipelinedForward` / `PipelinedHostForward` keep several such forwards in flight on one GPU.

`shard_batch` + `PeerGather` implement the multi-GPU partition of BASELINE.json's north_star: independent clouds of a
packed batch are split contiguously over the ranks and ONE all-gather of the per-cloud logits closes the step --
`PeerGather` does it with one own kernel over NVLink peer memory (csrc/peer_gather.cu) that is captured into the
forward's graph as an `epilogue`; `gather_logits` is the NCCL equivalent (fallback, unequal shards, CPU/gloo tests).
`global_box` and `gather_point_logits` serve the segmentation networks (batch-global bounding box for absolute radii,
per-point logits of unequal length).
"""
import torch

from . import runtime as _rt

_ALIGN = 256


def _layout(tensors):
    """Byte offsets of (points, batchIds, features) inside one packed block, each 256-byte aligned."""
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() * t.element_size() + _ALIGN - 1) // _ALIGN * _ALIGN
    return offs, max(total, _ALIGN)


def _views(block, like):
    offs, _ = _layout(like)
    out = []
    for o, t in zip(offs, like):
        nbytes = t.numel() * t.element_size()
        out.append(block[o:o + nbytes].view(t.dtype).view(t.shape))
    return tuple(out)


def pack_batch(pts, bids, feats, pin=False):
    """Pack a batch into the single-block layout the captured graph reads (same device as the inputs; pinned host
    memory if pin=True).  Returns a uint8 tensor."""
    _, total = _layout((pts, bids, feats))
    block = torch.empty((total,), dtype=torch.uint8, device=pts.device)
    if pin:
        block = block.pin_memory()
    for v, t in zip(_views(block, (pts, bids, feats)), (pts, bids, feats)):
        v.copy_(t)
    return block


class PlannedForward:
    def __init__(self, fn, out_rows, use_graph=True, slack=1.3, multistream=True, epilogue=None, exchange=None,
                 collect_in_graph=False, collect_deferred=False):
        """fn(points, batchIds, features) -> [>= out_rows, F] tensor; out_rows rows of it are the result.
        multistream: overlap the hierarchy / convolution-geometry / feature chains on three streams (runtime.planned).
        epilogue(result) -> tensor: applied to the out_rows result rows in every mode and CAPTURED INTO THE GRAPH with
        the forward -- the place for the closing all_gather of a sharded batch (NCCL collectives are capturable), so
        that a multi-GPU step still costs one graph launch on the host.
        exchange: a PeerExchange.  The `put` of the step's rows (+ the plan's overflow flag) becomes the last node of the
        graph; the matching `collect` is launched behind every run -- on `collect_stream` if set, so that this forward's
        stream never waits on a peer -- and leaves (rows of all ranks, meta) in `self.gathered`.
        collect_in_graph: capture the collect right behind the put instead (one graph launch per step and nothing else
        on the host: the better trade for the few-KB per-cloud logits of a sub-0.1-ms step; the side-stream collect is
        for per-point logits of multi-ms steps).
        collect_deferred: the collect of step k is issued right before the put of step k + 1 (same graph, same single
        host launch) and by flush() behind the last step: by then every peer's put of step k has long landed, so no
        forward ever waits on a peer -- ranks only couple through the exchange's ring of slots (4 steps of drift).
        `gathered` then holds step k's rows after the lane's next step or flush()."""
        self.fn = fn
        self.epilogue = epilogue
        self.exchange = exchange
        self.collect_in_graph = bool(collect_in_graph)
        self.collect_deferred = bool(collect_deferred)
        if self.collect_in_graph and self.collect_deferred:
            raise ValueError("PlannedForward: collect_in_graph and collect_deferred exclude each other")
        self.collect_stream = None
        self.gathered = None
        self.multistream = multistream
        self.out_rows = out_rows
        self.use_graph = use_graph
        self.slack = slack
        self.plan = None
        self.graph = None
        self._in_block = None     # static packed input block the graph reads
        self._static_in = None    # typed views of it
        self._result = None       # static float32 block: out_rows * F logits, then the overflow flag (int32 bits)
        self._out_view = None
        self._flag_view = None
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
        res = self._finish(out)
        self._collect()
        return res

    def _finish(self, out):
        res = out[: self.out_rows]
        if self.exchange is not None:
            res = res.contiguous()
            flag = self.plan.overflow if (self.plan is not None and _rt.mode() != "exact") else None
            if self.collect_deferred:
                self.gathered = self.exchange.collect(pending_only=True)   # the previous step's rows
            self.exchange.put(res, flag=flag)
            if self.collect_in_graph:
                self.gathered = self.exchange.collect()
        return self.epilogue(res) if self.epilogue is not None else res

    def _collect(self):
        """The collect that matches the put just enqueued (every put needs one on every rank)."""
        if self.exchange is None or self.collect_in_graph or self.collect_deferred:
            return
        cs = self.collect_stream
        if cs is None:
            self.gathered = self.exchange.collect()
            return
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self.gathered = self.exchange.collect()

    def flush(self):
        """collect_deferred: collect the last step's rows (waits for every rank's put of it) on the current stream."""
        if self.exchange is not None and self.collect_deferred:
            self.gathered = self.exchange.collect(pending_only=True)
        return self.gathered

    def _planned(self, pts, bids, feats):
        with _rt.planned(self.plan, pts.device, multistream=self.multistream):
            return self.fn(pts, bids, feats)

    def _planned_finish(self, pts, bids, feats):
        with _rt.planned(self.plan, pts.device, multistream=self.multistream):
            return self._finish(self.fn(pts, bids, feats))

    def _capture(self, pts, bids, feats):
        _, total = _layout((pts, bids, feats))
        self._in_block = torch.empty((total,), dtype=torch.uint8, device=pts.device)
        self._static_in = _views(self._in_block, (pts, bids, feats))
        for v, t in zip(self._static_in, (pts, bids, feats)):
            v.copy_(t)
        self.plan.begin(pts.device)
        side = torch.cuda.Stream(device=pts.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture (allocator, lazy module loading, NCCL communicator)
            self._planned_finish(*self._static_in)
            if self.exchange is not None and not self.collect_in_graph and not self.collect_deferred:
                self.exchange.collect()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.plan.overflow.zero_()
            out = self._planned_finish(*self._static_in)
            rows, cols = out.shape
            self._result = torch.empty((rows * cols + 1,), dtype=torch.float32, device=pts.device)
            self._out_view = self._result[: rows * cols].view(rows, cols)
            self._flag_view = self._result[rows * cols:].view(torch.int32)
            self._out_view.copy_(out)
            self._flag_view.copy_(self.plan.overflow)
        self.graph = g
        self.stats["captures"] += 1

    def _ready(self, pts, bids, feats):
        """True when a plan for these shapes exists (graph captured if use_graph)."""
        key = (tuple(pts.shape), tuple(feats.shape))
        if self.plan is None or key != self._shape_key:
            return False
        if self.plan.overflow is None:
            self.plan.begin(pts.device)
        if self.use_graph and self.graph is None:
            self._capture(pts, bids, feats)
        return True

    # -- public ------------------------------------------------------------------------------------------------
    def run_async(self, pts, bids, feats):
        """Enqueue one forward; returns (out[out_rows, F], overflow_flag int32[1]) device tensors, no host sync.
        The first call (and any call with new input shapes) records and therefore does synchronise."""
        if not self._ready(pts, bids, feats):
            out = self._record(pts, bids, feats)
            return out, torch.zeros((1,), dtype=torch.int32, device=pts.device)
        if not self.use_graph:
            self.plan.overflow.zero_()
            res = self._planned_finish(pts, bids, feats)
            self._collect()
            return res, self.plan.overflow
        for v, t in zip(self._static_in, (pts, bids, feats)):
            v.copy_(t, non_blocking=True)
        self.graph.replay()
        self._collect()
        self.stats["replays"] += 1
        return self._out_view, self._flag_view

    def run_packed(self, block):
        """Like run_async for a batch already packed with pack_batch (device or pinned host memory): one copy into
        the graph's input block, one graph launch.  Requires a captured graph (call run_async once before)."""
        self._in_block.copy_(block, non_blocking=True)
        self.graph.replay()
        self._collect()
        self.stats["replays"] += 1
        return self._out_view, self._flag_view

    def __call__(self, pts, bids, feats):
        """Forward with the overflow check (one D2H of the flag); exact re-run + re-plan when a capacity was hit."""
        out, flag = self.run_async(pts, bids, feats)
        if int(flag.item()) != 0:
            self.stats["overflows"] += 1
            self._no_solo_repair()
            out = self._record(pts, bids, feats)
        return out

    def _no_solo_repair(self):
        if self.epilogue is not None or self.exchange is not None:
            raise RuntimeError("mccnn_b200: a planned capacity was exceeded on a step whose graph contains a "
                               "collective; one rank cannot re-run it alone -- re-calibrate with a larger slack")

    def calibrate(self, batches):
        """Record several representative batches (same shapes) and plan for the element-wise maximum of their sizes."""
        sizes, flags = None, None
        for pts, bids, feats in batches:
            self._record(pts, bids, feats)
            s, f = self.plan.sizes, self.plan.flags
            sizes = list(s) if sizes is None else [max(a, b) for a, b in zip(sizes, s)]
            flags = list(f) if flags is None else [(a and b) if isinstance(a, bool) else b for a, b in zip(flags, f)]
        self.plan.sizes, self.plan.flags = sizes, flags
        self.plan.finalize()
        self.graph = None
        return self.plan


class HostForward:
    """End-to-end call for host-resident batches: pinned host block -> ONE H2D copy -> graph launch -> ONE D2H copy of
    logits + overflow flag -> host tensor.  This is the user-facing equivalent of the reference's
    `sess.run(logits, feed_dict)` (ModelNet/ModelNetEval.py:123-125)."""

    def __init__(self, engine, device):
        self.engine = engine
        self.device = torch.device(device)
        self._stage = None       # pinned packed staging for the three-tensor entry point
        self._result_host = None

    def _finish(self, unpack_for_rerun):
        eng = self.engine
        if self._result_host is None or self._result_host.numel() != eng._result.numel():
            self._result_host = torch.empty(eng._result.shape, dtype=torch.float32).pin_memory()
        self._result_host.copy_(eng._result, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        rows, cols = eng._out_view.shape
        out = self._result_host[: rows * cols].view(rows, cols)
        if int(self._result_host[rows * cols:].view(torch.int32)[0]) != 0:  # capacity exceeded: exact re-run, re-plan
            eng.stats["overflows"] += 1
            out = eng._record(*unpack_for_rerun()).cpu()
        return out

    def __call__(self, pts_host, bids_host, feats_host):
        """CPU tensors in (ideally pinned), pinned CPU logits [out_rows, F] out (buffer reused across calls)."""
        eng = self.engine
        key = (tuple(pts_host.shape), tuple(feats_host.shape))
        if eng.graph is None or eng._shape_key != key:  # first calls: record, then capture (through device tensors)
            return eng(*(t.to(self.device) for t in (pts_host, bids_host, feats_host))).cpu()
        _, total = _layout((pts_host, bids_host, feats_host))
        if self._stage is None or self._stage.numel() != total:
            self._stage = torch.empty((total,), dtype=torch.uint8).pin_memory()
        for v, t in zip(_views(self._stage, (pts_host, bids_host, feats_host)), (pts_host, bids_host, feats_host)):
            v.copy_(t)
        return self.call_packed(self._stage)

    def call_packed(self, pinned_block):
        """`pinned_block`: a batch packed with pack_batch(..., pin=True).  One H2D, one launch, one D2H."""
        eng = self.engine
        eng.run_packed(pinned_block)
        return self._finish(lambda: tuple(v.clone() for v in eng._static_in))

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

    submit(pts, bids, feats) / submit_packed(block) enqueue one forward on the next lane and return its ticket;
    result(ticket) makes the caller's stream (or the host) wait for that forward only and returns
    (out, overflow_flag) device tensors, valid until the lane is reused."""

    def __init__(self, fn, out_rows, depth=4, slack=1.3, multistream=True, epilogues=None, exchanges=None,
                 collect_in_graph=False, collect_deferred=False):
        """epilogues: one callable per lane (see PlannedForward).  Lanes run concurrently, so collectives in them must
        use ONE PROCESS GROUP PER LANE: an NCCL communicator may not be driven from two streams at once.
        exchanges: one PeerExchange per lane: the step's put is the last node of the lane's graph, its collect runs on
        a per-lane side stream (the lane's next forward never waits on a peer)."""
        self.depth = depth
        self.engines = [PlannedForward(fn, out_rows, use_graph=True, slack=slack, multistream=multistream,
                                       epilogue=None if epilogues is None else epilogues[i],
                                       exchange=None if exchanges is None else exchanges[i],
                                       collect_in_graph=collect_in_graph, collect_deferred=collect_deferred)
                        for i in range(depth)]
        self.streams = None
        self._events = [None] * depth
        self._after = [None] * depth
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
        for e in self.engines:
            if e.exchange is not None and not e.collect_in_graph and not e.collect_deferred:
                e.collect_stream = torch.cuda.Stream(device=dev)
        for e, s in zip(self.engines, self.streams):  # capture every lane's graph up front
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                e.run_async(*batches[0])
        torch.cuda.synchronize(dev)
        return plan

    def _enqueue(self, run, after=None):
        lane = self._next
        self._next = (self._next + 1) % self.depth
        s = self.streams[lane]
        s.wait_stream(torch.cuda.current_stream())  # inputs produced on the caller's stream
        with torch.cuda.stream(s):
            out, flag = run(self.engines[lane])
            # `after` runs ON THE LANE'S STREAM right behind the forward -- the place for the closing all_gather of a
            # sharded batch: issued from the caller's stream it would chain every lane behind the previous one
            self._after[lane] = after(out, flag) if after is not None else None
            ev = torch.cuda.Event()
            cs = self.engines[lane].collect_stream
            ev.record(s if cs is None else cs)  # the collect stream is ordered behind the forward
        self._events[lane] = ev
        return lane

    def submit(self, pts, bids, feats, after=None):
        return self._enqueue(lambda e: e.run_async(pts, bids, feats), after)

    def submit_packed(self, block, after=None):
        return self._enqueue(lambda e: e.run_packed(block), after)

    def after_result(self, ticket):
        """What the `after` callback of that submission returned (valid once result()/join() made it visible)."""
        return self._after[ticket]

    def result(self, ticket, wait_on_host=False):
        ev = self._events[ticket]
        if wait_on_host:
            ev.synchronize()
        else:
            torch.cuda.current_stream().wait_event(ev)
        e = self.engines[ticket]
        return e._out_view, e._flag_view

    def gathered(self, ticket):
        """(rows of all ranks, meta) of that submission's exchange (valid once result()/join() made it visible)."""
        return self.engines[ticket].gathered

    def join(self):
        """The caller's stream waits for every lane; deferred collects are flushed first (every put is collected)."""
        for lane, e in enumerate(self.engines):
            if e.exchange is not None and e.collect_deferred and self._events[lane] is not None:
                with torch.cuda.stream(self.streams[lane]):
                    e.flush()
                    ev = torch.cuda.Event()
                    ev.record(self.streams[lane])
                self._events[lane] = ev
        for ev in self._events:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)


class PipelinedHostForward:
    """End-to-end throughput mode for host-resident batches: every lane of a PipelinedForward has its own pinned result
    buffer, so the H2D copy of batch i+1, the forward of batch i and the D2H of batch i-1 overlap.
    submit_packed(pinned_block) returns a ticket; result(ticket) blocks until that batch's logits are on the host."""

    def __init__(self, pipe, device):
        self.pipe = pipe
        self.device = torch.device(device)
        self._result_host = [None] * pipe.depth
        self._extra_host = [None] * pipe.depth
        self._has_extra = [None] * pipe.depth
        self._events = [None] * pipe.depth
        self._blocks = [None] * pipe.depth

    def submit_packed(self, pinned_block, after=None):
        """after(out) -> device tensor (e.g. engine.gather_logits over the ranks), evaluated on the lane's stream behind
        the forward; result() then returns ITS host copy instead of the local logits.  Lanes with a PeerExchange return
        the gathered rows of all ranks."""
        lane = self.pipe._next
        if self._blocks[lane] is not None:
            raise RuntimeError("PipelinedHostForward: %d results outstanding -- call result() on the oldest ticket "
                               "before submitting another batch" % self.pipe.depth)
        eng, s = self.pipe.engines[lane], self.pipe.streams[lane]
        if self._result_host[lane] is None:
            self._result_host[lane] = torch.empty(eng._result.shape, dtype=torch.float32).pin_memory()
        self.pipe._next = (lane + 1) % self.pipe.depth
        self._has_extra[lane] = None
        with torch.cuda.stream(s):
            eng.run_packed(pinned_block)                                   # one H2D + one graph launch (+ collect)
            eng.flush()                                                    # deferred collect: the host wants THIS step
            last = s
            if eng.exchange is not None:                # ONE D2H: rows of all ranks + (rows, flag) of every rank
                last = eng.collect_stream if eng.collect_stream is not None else s
                blk = eng.exchange.block
                if self._extra_host[lane] is None:
                    self._extra_host[lane] = torch.empty(blk.shape, dtype=blk.dtype).pin_memory()
                with torch.cuda.stream(last):
                    self._extra_host[lane].copy_(blk, non_blocking=True)
                self._has_extra[lane] = "exchange"
            else:
                self._result_host[lane].copy_(eng._result, non_blocking=True)  # one D2H: logits + overflow flag
                if after is not None:
                    extra = after(eng._out_view)
                    if self._extra_host[lane] is None or self._extra_host[lane].shape != extra.shape:
                        self._extra_host[lane] = torch.empty(extra.shape, dtype=extra.dtype).pin_memory()
                    self._extra_host[lane].copy_(extra, non_blocking=True)
                    extra.record_stream(s)
                    self._has_extra[lane] = "after"
            ev = torch.cuda.Event()
            ev.record(last)
        self._events[lane], self._blocks[lane] = ev, pinned_block
        return lane

    def submit(self, pts_host, bids_host, feats_host):
        return self.submit_packed(pack_batch(pts_host, bids_host, feats_host, pin=True))

    def result(self, ticket):
        self._events[ticket].synchronize()
        self._blocks[ticket] = None
        eng = self.pipe.engines[ticket]
        rows, cols = eng._out_view.shape
        host = self._result_host[ticket]
        if self._has_extra[ticket] == "exchange":
            ex, blk = eng.exchange, self._extra_host[ticket]
            n = ex.world * ex.cap_rows * ex.cols
            out, meta = blk[:n].view(ex.world * ex.cap_rows, ex.cols), blk[n:].view(torch.int32)
            world = ex.world
            if int(meta[2 * world + 1]) != 0:
                # the flags travelled with the rows: EVERY rank sees the overflow of ANY rank and stops here -- no rank
                # returns another rank's truncated rows, and no rank re-records alone
                eng.stats["overflows"] += 1
                raise RuntimeError("mccnn_b200: a planned capacity was exceeded on rank(s) %s of a sharded step; "
                                   "re-calibrate with a larger slack" %
                                   [r for r in range(world) if int(meta[2 * r + 1]) != 0])
            return out[: int(meta[2 * world])] if eng.exchange.compact else out
        if int(host[rows * cols:].view(torch.int32)[0]) != 0:  # planned capacity exceeded: exact re-run of this batch
            eng.stats["overflows"] += 1
            eng._no_solo_repair()
            with torch.cuda.stream(self.pipe.streams[ticket]):
                out = eng._record(*(v.clone() for v in eng._static_in))
                eng.run_async(*(v.clone() for v in eng._static_in))  # re-capture for the following steps
            self.pipe.streams[ticket].synchronize()
            return out.cpu()
        if self._has_extra[ticket] == "after":
            return self._extra_host[ticket]
        return host[: rows * cols].view(rows, cols)


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
    if batch_size % world_size == 0 and local_logits.shape[0] * world_size == batch_size:
        # equal shards (the benchmark's case): the collective alone, no padding / re-assembly kernels
        out = local_logits.new_empty((batch_size, local_logits.shape[1]))
        dist.all_gather_into_tensor(out, local_logits.contiguous(), group=group)
        return out
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


class PeerGather:
    """gather_logits for equal shards over NVLink PEER MEMORY instead of NCCL (csrc/peer_gather.cu): every rank stores
    its [rows, cols] logits straight into a symmetric buffer on every peer and waits on flag words -- one tiny kernel
    that is captured into the forward's CUDA graph like any other node.  One instance per concurrent lane.

    The symmetric buffers come from torch.distributed._symmetric_memory (CUDA VMM allocations exported to the peers of
    `group`); construction is collective.  Raises if symmetric memory is unavailable (caller falls back to NCCL)."""

    def __init__(self, rows, cols, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from ._lib import lib
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows, self.cols, self.n = rows, cols, rows * cols
        floats = int(lib.mc_allgather_rows_p2p_buffer_floats(self.n, self.world))
        self.buf = symm.empty((floats,), dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)            # my flags are zero before any peer can learn my address
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer addresses")
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.counter = torch.zeros((1,), dtype=torch.int32, device=device)
        dist.barrier(group)

    def __call__(self, local_logits):
        from . import ops
        if tuple(local_logits.shape) != (self.rows, self.cols):
            raise ValueError("PeerGather: expected logits of shape %s" % ((self.rows, self.cols),))
        out = local_logits.new_empty((self.world * self.rows, self.cols))
        ops.allgather_rows_p2p(local_logits, self.ptrs, self.counter, self.rank, self.world, out)
        return out


class PeerExchange:
    """The closing all-gather of a sharded step as a put / collect pair over NVLink peer memory (csrc/peer_gather.cu),
    for per-cloud logits (equal shards) and per-point logits (unequal row counts, `compact=True`) alike.

      put(rows, n_dev=None, flag=None)   enqueue on the current stream (capturable: the last node of a forward's graph):
                                         this rank's live rows + its overflow flag -> every rank's symmetric buffer;
      collect() -> (out, meta)           enqueue on the current stream: waits for every rank's put of the step.
                                         out [world * cap_rows, cols]: rank r's rows at r * cap_rows, or (compact) all
                                         live rows contiguous in rank order == the order of the unsharded batch;
                                         meta int32 [2 * world + 2]: (rows, flag) per rank, total rows, MAX of the flags.

    Every put must be matched by one collect on every rank; ranks may drift apart by `slots` steps before a put blocks.
    One instance per concurrent lane.  Construction is collective (symmetric-memory rendezvous over `group`)."""

    def __init__(self, cap_rows, cols, device, slots=4, compact=False, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from ._lib import lib
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cap_rows, self.cols, self.slots, self.compact = int(cap_rows), int(cols), int(slots), bool(compact)
        floats = int(lib.mc_peer_exchange_buffer_floats(self.cap_rows, self.cols, self.world, self.slots))
        self.buf = symm.empty((floats,), dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)            # my flags / acks are zero before any peer can learn my address
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer addresses")
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.state = torch.zeros((4,), dtype=torch.int32, device=device)
        # gathered rows and meta share one block: a host consumer needs ONE device-to-host copy per step
        n = self.world * self.cap_rows * self.cols
        self.block = torch.zeros((n + 2 * self.world + 2,), dtype=torch.float32, device=device)
        self.out = self.block[:n].view(self.world * self.cap_rows, self.cols)
        self.meta = self.block[n:].view(torch.int32)
        dist.barrier(group)

    def put(self, rows, n_dev=None, flag=None):
        from . import ops
        if rows.dim() != 2 or rows.shape[1] != self.cols or rows.shape[0] > self.cap_rows:
            raise ValueError("PeerExchange.put: expected at most %d rows of %d columns" % (self.cap_rows, self.cols))
        if rows.shape[0] < self.cap_rows and n_dev is None:
            n_dev = torch.full((1,), rows.shape[0], dtype=torch.int32, device=rows.device)
        ops.peer_put(rows, self.ptrs, self.state, self.rank, self.world, self.slots, self.cap_rows, n_dev, flag)

    def collect(self, pending_only=False):
        """pending_only: a no-op unless a put of this exchange is still uncollected (decided on the device, so the call
        can sit in a CUDA graph that runs both as a lane's first step and as a later one)."""
        from . import ops
        return ops.peer_collect(self.ptrs, self.state, self.rank, self.world, self.slots, self.cap_rows, self.cols,
                                self.out, self.meta, self.compact, pending_only=pending_only)

    def __call__(self, rows):
        """put + collect back to back on the current stream (gathered rows inside the same graph)."""
        self.put(rows)
        return self.collect()[0]


def gather_point_logits(local_logits, world_size, group=None):
    """Segmentation variant of the closing collective: per-point logits [N_r, F] with a
    different N_r on every rank: an 8-byte all_gather of the row counts, then ONE all_gather of blocks padded to the
    largest N_r.  Returns [sum N_r, F] in rank order == the order of the unsharded packed batch."""
    import torch.distributed as dist
    if world_size == 1:
        return local_logits
    n, f = local_logits.shape
    counts = torch.zeros((world_size,), dtype=torch.int64, device=local_logits.device)
    mine = torch.tensor([n], dtype=torch.int64, device=local_logits.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    rows = int(counts.max().item())
    padded = local_logits.new_zeros((rows, f))
    padded[:n] = local_logits
    out = local_logits.new_empty((world_size * rows, f))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * rows: r * rows + int(counts[r].item())] for r in range(world_size)], dim=0)


def global_box(aabb_min, aabb_max, group=None):
    """Batch-global bounding box across the ranks for `relativeRadius=False` networks (MCSegScanNet): the reference
    takes ONE box over all points of the batch (tf_ops/aabb_gpu.cu:104-114) and derives the grid origin and the number
    of cells from it (tf_ops/sort_gpu.cu:383-389), so shards must agree on it to reproduce the unsharded cell phases
    and hence the same Poisson samples.  ONE 6-float collective: min and max are fused into a single MIN all-reduce
    by negating the max; every row of the [B_local, 3] outputs holds the merged box (as every row of the reference's
    output does in this mode)."""
    import torch.distributed as dist
    both = torch.cat([aabb_min[0], -aabb_max[0]]).contiguous()       # 6 floats: ranks may hold different row counts
    dist.all_reduce(both, op=dist.ReduceOp.MIN, group=group)
    rows = aabb_min.shape[0]
    return both[:3].expand(rows, 3).contiguous(), (-both[3:]).expand(rows, 3).contiguous()
