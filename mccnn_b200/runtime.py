"""Execution modes of the op API: exact shapes (drop-in, syncs like the reference) vs planned capacities (sync-free).

The reference's ops have two data-dependent output sizes -- the number of Poisson samples M
(tf_ops/poisson_sampling.cu:222) and the number of neighbour pairs K (tf_ops/find_neighbors.cu:308) -- and blocks the
host on each of them, >= 6 times per MCClassS forward.  Here:

  * exact mode (default): the op wrappers read the count (one D2H each) and return exactly-shaped tensors, so any code
    written against the reference API works unchanged;
  * `with record() as plan:` -- exact mode that also logs every dynamic size in call order;
  * `with planned(plan):`   -- the same Python code runs again, but every dynamic output is allocated with the planned
    capacity (recorded size x slack) and carries its live row count as a device int (`tensor._mc_count`); the kernels
    read the counts on the device (C-ABI `(capacity, dCount)` convention), nothing blocks, and the whole forward can
    be captured in a CUDA graph.  If a live size ever exceeds its capacity a device flag (`plan.overflow`) is
    raised; the caller checks it with the result and re-records.
"""
import contextlib

import torch


class Plan:
    def __init__(self, slack=1.3, granule=256):
        self.sizes = []        # recorded live sizes in call order
        self.capacities = []   # planned capacities, same order
        self.flags = []        # recorded boolean decisions in call order (e.g. "this batch is cloud-contiguous")
        self._flag_cursor = 0
        self.slack = slack
        self.granule = granule
        self.overflow = None   # int32[1] device flag, set by kernels when a capacity is exceeded
        self._cursor = 0

    def finalize(self):
        g = self.granule
        self.capacities = [max(g, (int(s * self.slack) + g - 1) // g * g) for s in self.sizes]
        return self

    def begin(self, device):
        self._cursor = 0
        self._flag_cursor = 0
        if self.overflow is None or self.overflow.device != device:
            self.overflow = torch.zeros((1,), dtype=torch.int32, device=device)

    def next_flag(self):
        if self._flag_cursor >= len(self.flags):
            raise RuntimeError("mccnn_b200: the planned run takes more recorded decisions than were recorded")
        f = self.flags[self._flag_cursor]
        self._flag_cursor += 1
        return f

    def next_capacity(self):
        if self._cursor >= len(self.capacities):
            raise RuntimeError("mccnn_b200: the planned run issues more dynamic ops than were recorded")
        c = self.capacities[self._cursor]
        self.last_recorded = self.sizes[self._cursor]  # live size seen when recording: a host-side hint
        self._cursor += 1
        return c


class _State:
    mode = "exact"   # "exact" | "record" | "planned"
    plan = None
    lane = "main"    # "main" (features, dense glue) | "hier" (point hierarchy) | "geom" (convolution geometry)
    streams = None   # {lane: torch.cuda.Stream} while a multi-stream planned run is active
    forked = None    # side streams that joined the run so far
    start_event = None
    shard = None     # (process group, world size) while the batch is one shard of a multi-GPU batch


_state = _State()
_side_streams = {}   # (device index, lane) -> torch.cuda.Stream, created once


def mode():
    return _state.mode


def current_plan():
    return _state.plan


@contextlib.contextmanager
def record(slack=1.3):
    prev = (_state.mode, _state.plan)
    plan = Plan(slack=slack)
    _state.mode, _state.plan = "record", plan
    try:
        yield plan
    finally:
        _state.mode, _state.plan = prev
        plan.finalize()


@contextlib.contextmanager
def planned(plan, device=None, multistream=True):
    """Sync-free execution against `plan`.  With multistream, the three independent chains of a forward --
    hierarchy (sort + Poisson sampling per level), convolution geometry (grid, neighbour list, pdf) and features
    (feature re-sort, spatial_conv, dense glue) -- are issued on three streams linked by events, so they overlap on
    the GPU; captured in a CUDA graph this becomes a DAG instead of a chain.  All side streams are joined back into
    the caller's stream when the context exits."""
    prev = (_state.mode, _state.plan, _state.streams, _state.forked, _state.start_event)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    plan.begin(device)
    _state.mode, _state.plan = "planned", plan
    if multistream:
        streams = {}
        for ln in ("hier", "geom"):
            key = (device.index, ln)
            if key not in _side_streams:
                _side_streams[key] = torch.cuda.Stream(device=device)
            streams[ln] = _side_streams[key]
        _state.streams, _state.forked = streams, []
        _state.start_event = torch.cuda.Event()
        _state.start_event.record(torch.cuda.current_stream(device))
    else:
        _state.streams, _state.forked, _state.start_event = None, None, None
    try:
        yield plan
    finally:
        if _state.forked:
            main = torch.cuda.current_stream(device)
            for s in _state.forked:  # join
                ev = torch.cuda.Event()
                ev.record(s)
                main.wait_event(ev)
        _state.mode, _state.plan, _state.streams, _state.forked, _state.start_event = prev


@contextlib.contextmanager
def lane(name):
    """Declares which chain the ops issued inside belong to (used by MCConvBuilder)."""
    prev = _state.lane
    _state.lane = name
    try:
        yield
    finally:
        _state.lane = prev


def _tensors(out):
    if isinstance(out, torch.Tensor):
        return [out]
    if isinstance(out, (tuple, list)):
        return [o for o in out if isinstance(o, torch.Tensor)]
    return []


def dispatch(inputs, fn):
    """Run fn() on the stream of the current lane.  Outside multi-stream planned mode this is fn().  Otherwise the
    stream first waits for the readiness events of the tensor inputs produced on other streams, and the outputs are
    tagged with an event recorded after fn() so that consumers on other streams can wait for them."""
    if _state.streams is None:
        return fn()
    main = torch.cuda.current_stream()
    s = _state.streams.get(_state.lane)
    if s is None:  # main lane: wait for side-stream producers
        for t in inputs:
            ev = getattr(t, "_mc_ev", None)
            if ev is not None:
                main.wait_event(ev)
                if t.is_cuda:
                    t.record_stream(main)
        return fn()
    if s not in _state.forked:  # fork: the side stream starts after everything already queued on the main stream
        s.wait_event(_state.start_event)
        _state.forked.append(s)
    with torch.cuda.stream(s):
        for t in inputs:
            ev = getattr(t, "_mc_ev", None)
            if ev is not None and getattr(t, "_mc_stream", None) is not s:
                s.wait_event(ev)
            if t.is_cuda:
                t.record_stream(s)
        out = fn()
        ev = torch.cuda.Event()
        ev.record(s)
    for o in _tensors(out):
        o._mc_ev = ev
        o._mc_stream = s
    return out


@contextlib.contextmanager
def sharded(group=None, world_size=None):
    """Declares that the batch being processed is one shard of a larger batch spread over the ranks of `group`
    (torch.distributed; None = the default group).  The only op sequence that needs to know is a hierarchy with
    absolute radii, whose batch-global bounding box is then all-reduced (engine.global_box)."""
    import torch.distributed as dist
    prev = _state.shard
    ws = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    _state.shard = (group, ws) if ws > 1 else None
    try:
        yield
    finally:
        _state.shard = prev


def shard():
    """(group, world_size) while inside `sharded(...)` with more than one rank, else None."""
    return _state.shard


def count_of(t):
    """Live row count (int32[1] device tensor) attached to a capacity-sized tensor, or None for exact tensors."""
    return getattr(t, "_mc_count", None)


def tag(t, count, hint=None):
    """Attach the live row count (device int) and, optionally, a host-side estimate of it (the size recorded when the
    plan was made; used only to pick kernel variants / launch shapes, never for correctness)."""
    if count is not None:
        t._mc_count = count
    if hint is not None:
        t._mc_hint = hint
    return t


def hint_of(t):
    return getattr(t, "_mc_hint", None)


class DeferredAffine:
    """act(base * scale + shift), not computed yet: what an inference batch norm (+ ReLU) returns.

    Every consumer inside this package folds the pending per-column affine into its own kernel -- the prologue of the
    next dense layer, or the row permutation that brings the features of a convolution into grid order -- so the
    batch-norm layers of a network cost no launch and no pass over the features.  For anything else the object
    behaves like the tensor it stands for: torch functions, attribute access and indexing materialise it (one
    scale_shift_act launch, cached)."""

    def __init__(self, base, scale, shift, relu):
        self.base, self.scale, self.shift, self.relu = base, scale, shift, bool(relu)
        self._value = None

    shape = property(lambda self: self.base.shape)
    dtype = property(lambda self: self.base.dtype)
    device = property(lambda self: self.base.device)
    is_cuda = property(lambda self: self.base.is_cuda)
    _mc_count = property(lambda self: count_of(self.base))
    _mc_hint = property(lambda self: hint_of(self.base))

    def dim(self):
        return self.base.dim()

    def materialize(self):
        if self._value is None:
            from . import ops
            nd = count_of(self.base)
            self._value = dispatch([self.base], lambda: tag(
                ops.scale_shift_act(self.base, self.scale, self.shift, self.relu, nd), nd, hint_of(self.base)))
        return self._value

    def __getattr__(self, name):  # only reached for names not defined above
        if name.startswith("__") or name in ("base", "_value"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, key):
        return self.materialize()[key]

    def __len__(self):
        return self.base.shape[0]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def real(a):
            if isinstance(a, DeferredAffine):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(real(x) for x in a)
            return a
        return func(*real(args), **{k: real(v) for k, v in (kwargs or {}).items()})


def materialize(t):
    return t.materialize() if isinstance(t, DeferredAffine) else t
