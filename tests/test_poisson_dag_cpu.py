"""The reference's Poisson-disk sampling is a priority rule, not 27 data-dependent passes.

tf_ops/poisson_sampling.cu:175-230 walks 27 colour phases of the grid, cell by cell, point by point, accepting a point
iff no ALREADY ACCEPTED point lies within r in its 27 cells.  Equivalent statement: a point is accepted iff no accepted
point of EARLIER PRIORITY lies within r, with priority = (phase of the cell, then position inside the cell) -- cells of
one phase are three cells apart and never interact.  That makes the accepted set the unique fixpoint of a relaxation
over the "can block" DAG, which is what the round-2 GPU experiment (profiles/
r2_poisson_dependency_graph_experiment.txt) evaluated in parallel.  This test states the equivalence on the CPU oracle:
a NumPy relaxation over the DAG selects exactly the oracle's samples."""
import numpy as np
import pytest

from common import make_batch
from oracle import cpu as O

# phase (position in the reference's shuffled offset list, poisson_sampling.cu:192-196) of the cell offset
# (x % 3 - 1, y % 3 - 1, z % 3 - 1), indexed (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1) -- as csrc/poisson.cu holds it
PHASE_OF_OFFSET = [25, 15, 26, 24, 12, 22, 6, 16, 11, 7, 5, 1, 18, 17, 4, 14, 3, 2, 13, 21, 9, 20, 8, 10, 0, 19, 23]


def _dag_select(sp, sb, cells, mn, mx, radius, scale_inv):
    B, n = cells.shape[0], cells.shape[1]
    N = sp.shape[0]
    cell_of = np.full(N, -1, np.int64)
    phase_of = np.zeros(N, np.int64)
    for b in range(B):
        for x in range(n):
            for y in range(n):
                for z in range(n):
                    lo, hi = cells[b, x, y, z]
                    if hi > lo:
                        cell_of[lo:hi] = ((b * n + x) * n + y) * n + z
                        phase_of[lo:hi] = PHASE_OF_OFFSET[(x % 3) * 9 + (y % 3) * 3 + (z % 3)]
    ext = np.max(mx.astype(np.float32) - mn.astype(np.float32), axis=1)
    preds, margin = [[] for _ in range(N)], np.inf
    for i in range(N):
        c = cell_of[i]
        b, rem = divmod(c, n ** 3)
        x, rem = divmod(rem, n * n)
        y, z = divmod(rem, n)
        r = float(np.float32(radius) * ext[b]) if scale_inv else float(np.float32(radius))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    qx, qy, qz = x + dx, y + dy, z + dz
                    if not (0 <= qx < n and 0 <= qy < n and 0 <= qz < n):
                        continue
                    lo, hi = cells[b, qx, qy, qz]
                    for j in range(lo, hi):
                        earlier = (j < i) if cell_of[j] == c else (phase_of[j] < phase_of[i])
                        if not earlier:
                            continue
                        d = float(np.linalg.norm(sp[j].astype(np.float64) - sp[i].astype(np.float64)))
                        margin = min(margin, abs(d - r) / r)
                        if d < r:
                            preds[i].append(j)
    status = np.zeros(N, np.int8)                     # 0 undecided, 1 accepted, 2 rejected
    rounds = 0
    while (status == 0).any():
        rounds += 1
        new = status.copy()
        for i in np.nonzero(status == 0)[0]:
            st = status[preds[i]] if preds[i] else np.zeros(0, np.int8)
            if (st == 1).any():
                new[i] = 2
            elif (st != 0).all():
                new[i] = 1
        assert not np.array_equal(new, status), "the relaxation must make progress: the graph is acyclic"
        status = new
    return np.nonzero(status == 1)[0], phase_of, rounds, margin


@pytest.mark.parametrize("B,N,dist,radius,scale_inv", [(2, 300, "sphere", 0.1, True), (3, 200, "cube", 0.15, True),
                                                        (2, 400, "room", 0.4, False), (1, 150, "sphere", 0.45, True)])
def test_priority_dag_fixpoint_is_the_reference_sample_set(B, N, dist, radius, scale_inv):
    pts, bids, feats = make_batch(B, N, dist, seed=77, num_features=1)
    mn, mx = O.compute_aabb(pts, bids, B, scale_inv)
    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, radius, scale_inv)
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, radius, scale_inv)
    op, ob, oi, oph = O.poisson_sampling(sp, sb, cells, mn, mx, radius, B, scale_inv, return_phase=True)
    accepted, phase_of, rounds, margin = _dag_select(sp, sb, cells, mn, mx, radius, scale_inv)
    assert margin > 1e-6          # no pair sits on the float32 rounding edge of the predicate (float64 used here)
    assert np.array_equal(np.sort(oi), accepted)
    assert np.array_equal(phase_of[oi], oph)                  # and the phases are the oracle's
    assert rounds <= 27 * 4       # the serial depth the relaxation needed: far below "27 phases x points per cell"
