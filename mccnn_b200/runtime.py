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
        self._cursor += 1
        return c


class _State:
    mode = "exact"   # "exact" | "record" | "planned"
    plan = None


_state = _State()


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
def planned(plan, device=None):
    prev = (_state.mode, _state.plan)
    plan.begin(torch.device("cuda", torch.cuda.current_device()) if device is None else device)
    _state.mode, _state.plan = "planned", plan
    try:
        yield plan
    finally:
        _state.mode, _state.plan = prev


def count_of(t):
    """Live row count (int32[1] device tensor) attached to a capacity-sized tensor, or None for exact tensors."""
    return getattr(t, "_mc_count", None)


def tag(t, count):
    if count is not None:
        t._mc_count = count
    return t
