"""Per-op CUDA-event timers with algorithmic byte counts (the reference prints one wall-clock number per forward,
ModelNet/ModelNetEval.py:141; this is the per-op view of the same forward).

    with OpProfile(batch_size) as prof:
        forward(...)                     # any code that goes through mccnn_b200.ops (exact or planned mode)
    rows = prof.table()                  # synchronises, then: [{op, call, us, alg_bytes, gbs}, ...] in call order

Every call of an op of the C-ABI layer (mccnn_b200.ops) is bracketed by a CUDA-event pair on the stream it is issued
on; its ALGORITHMIC bytes come from SURVEY.md section 8(d)'s per-op formulas evaluated on the LIVE sizes of that call
(device-side counts are read after the forward, so profiling adds no synchronisation to the timed ops).
bench.py attaches these tables to its JSON line (`kernels`, `other_configs[*].ops`).
"""
import collections

import torch

from . import ops, runtime


def _live(t):
    c = runtime.count_of(t)
    return int(c.item()) if c is not None else int(t.shape[0])


def _cells(t):
    return int(t.numel() // 2)


def _count(t):
    return int(t.item())


def _dense_bytes(rows, k, n):
    return 4 * (rows * k + k * n + rows * n)


def _rows(x, kw):
    nd = kw.get("n_dev")
    return int(nd.item()) if nd is not None else int(x.shape[0])


# op -> bytes(args, kwargs, result, batch_size); SURVEY.md section 8(d), 4-byte elements
ALGORITHMIC_BYTES = {
    "compute_aabb": lambda a, kw, r, B: 16 * a[0].shape[0] + 24 * B,
    "build_grid": lambda a, kw, r, B: _live(a[0]) * (48 + 8 * a[2].shape[1]) + 20 * _cells(r[5]),
    "sort_points_step1": lambda a, kw, r, B: _live(a[0]) * 24 + 12 * B * a[5] ** 3,
    "sort_points_step2": lambda a, kw, r, B: _live(a[0]) * (24 + 8 * a[2].shape[1]) + 8 * _cells(r[3]),
    "sort_features": lambda a, kw, r, B: _live(a[0]) * (8 * a[0].shape[1] + 4),
    "sort_features_back": lambda a, kw, r, B: _live(a[0]) * (8 * a[0].shape[1] + 4),
    "get_sampled_features": lambda a, kw, r, B: _live(a[0]) * (4 + 8 * a[1].shape[1]),
    "transform_indexs": lambda a, kw, r, B: 8 * _live(a[1]) + 8 * _live(a[0]),
    "find_neighbors_fused": lambda a, kw, r, B: (16 * _live(a[0]) + 12 * _live(a[2]) +
                                                 8 * min(_cells(a[3]), 27 * _live(a[0])) + 4 * _live(a[0]) +
                                                 8 * _count(r[2])),
    "find_neighbors_pdf_fused": lambda a, kw, r, B: (16 * _live(a[0]) + 12 * _live(a[2]) +
                                                     8 * min(_cells(a[3]), 27 * _live(a[0])) + 4 * _live(a[0]) +
                                                     12 * _count(r[3]) + 12 * _live(a[2])),
    "find_neighbors_count": lambda a, kw, r, B: (16 * _live(a[0]) + 12 * _live(a[2]) +
                                                 8 * min(_cells(a[3]), 27 * _live(a[0])) + 4 * _live(a[0])),
    "find_neighbors_pack": lambda a, kw, r, B: (16 * _live(a[0]) + 12 * _live(a[2]) +
                                                8 * min(_cells(a[3]), 27 * _live(a[0])) + 4 * _live(a[0]) +
                                                8 * _live(r)),
    "compute_pdf": lambda a, kw, r, B: 12 * _live(a[5]) + 4 * _live(a[4]) + 12 * _live(a[0]),
    "poisson_sampling": lambda a, kw, r, B: 13 * _live(a[0]) + 8 * _cells(a[2]) + 20 * _count(r[3]),
    "spatial_conv": lambda a, kw, r, B: (12 * _live(a[6]) + 4 * a[1].shape[1] * _live(a[6]) + 12 * _live(a[0]) +
                                         16 * _live(a[4]) + 4 * r.shape[1] * _live(a[4]) +
                                         4 * sum(int(w.numel()) for w in a[9:15])),
    "scale_shift_act": lambda a, kw, r, B: 8 * _rows(a[0], kw) * a[0].shape[1],
    "dense_layer": lambda a, kw, r, B: _dense_bytes(_rows(a[0], kw), a[1].shape[0], a[1].shape[1]),
    "dense_layer_fewrows": lambda a, kw, r, B: _dense_bytes(_rows(a[0], kw), a[1].shape[0], a[1].shape[1]),
    "dense_layer_tc": lambda a, kw, r, B: _dense_bytes(_rows(a[0], kw), a[1].shape[1], a[1].shape[0]),
    "mlp_rows": lambda a, kw, r, B: sum(_dense_bytes(_rows(a[0], kw), l[0].shape[0], l[0].shape[1]) for l in a[1]),
}


class OpProfile:
    def __init__(self, batch_size, names=None):
        self.B = batch_size
        self.names = [n for n in (names or ALGORITHMIC_BYTES) if hasattr(ops, n)]
        self.records = []
        self._orig = {}

    def _wrap(self, name, fn):
        def timed(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            self.records.append((name, s, e, a, kw, r))
            return r
        return timed

    def __enter__(self):
        for n in self.names:
            self._orig[n] = getattr(ops, n)
            setattr(ops, n, self._wrap(n, self._orig[n]))
        return self

    def __exit__(self, *exc):
        for n, fn in self._orig.items():
            setattr(ops, n, fn)
        self._orig = {}
        return False

    def table(self):
        """One row per op call, in call order (synchronises the device)."""
        torch.cuda.synchronize()
        calls = collections.Counter()
        rows = []
        for name, s, e, a, kw, r in self.records:
            us = s.elapsed_time(e) * 1e3
            by = int(ALGORITHMIC_BYTES[name](a, kw, r, self.B)) if name in ALGORITHMIC_BYTES else 0
            rows.append({"op": name, "call": calls[name], "us": us, "alg_bytes": by,
                         "gbs": by / us * 1e-3 if us > 0 else None})
            calls[name] += 1
        return rows

    @staticmethod
    def by_op(rows):
        """Aggregate table(): op -> {calls, us, alg_bytes, gbs}, sorted by time."""
        agg = collections.OrderedDict()
        for r in rows:
            d = agg.setdefault(r["op"], {"calls": 0, "us": 0.0, "alg_bytes": 0})
            d["calls"] += 1
            d["us"] += r["us"]
            d["alg_bytes"] += r["alg_bytes"]
        for d in agg.values():
            d["gbs"] = d["alg_bytes"] / d["us"] * 1e-3 if d["us"] > 0 else None
        return collections.OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["us"]))
