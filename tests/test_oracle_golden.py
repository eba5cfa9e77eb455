"""Pins the CPU oracle (oracle/mccnn_oracle.c) to outputs of the REFERENCE's own CUDA kernels.

tests/golden/*.npz were produced on a B200 by tests/golden/make_golden.py from oracle/_ref/libmccnn_ref.so, i.e. the
unmodified /root/reference/tf_ops/*.cu compiled for sm_100a.  The reference ships no tests or golden vectors of its
own; this is the pin.  Runs without a GPU.

Bit-exact: aabb, keys, cell table, neighbour list, Poisson accepted set, gathers.  1e-5 relative: pdf, convolution.
The reference's sort order inside a cell and its Poisson output order inside a (cloud, phase) launch are arbitrary
(atomic arrival order, tf_ops/sort_gpu.cu:170, tf_ops/poisson_sampling.cu:115): those are compared as permutations.
"""
import glob
import os

import numpy as np
import pytest

from common import make_conv_weights, rel_err, sum_err
from oracle import cpu as O

FILES = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
               if not os.path.basename(f).startswith(("model_", "dataset_")))  # those: tests/test_models_*.py, test_dataset_*.py


def test_fixtures_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_matches_reference_kernels(path):
    g = np.load(path)
    B, rg, rs, win, C, Co, combin, sinv, avg = g["params"]
    B, C, Co, combin, sinv, avg = int(B), int(C), int(Co), bool(combin), bool(sinv), bool(avg)
    pts, bids, feats = g["points"], g["batch_ids"], g["features"]
    N = pts.shape[0]

    mn, mx = O.compute_aabb(pts, bids, B, sinv)
    assert np.array_equal(mn, g["aabb_min"]) and np.array_equal(mx, g["aabb_max"])

    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, rg, sinv)
    assert np.array_equal(keys, g["ref_keys"])
    ridx = g["ref_indexs_raw"]
    assert np.array_equal(np.sort(ridx), np.arange(N))                   # the reference's result is a permutation
    rkeys_sorted = np.empty_like(keys)
    rkeys_sorted[ridx] = keys
    okeys_sorted = np.empty_like(keys)
    okeys_sorted[idx] = keys
    assert np.array_equal(rkeys_sorted, okeys_sorted)                    # ... into the same cells
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, rg, sinv)
    assert np.array_equal(cells, g["ref_cell_indexs"])
    flat = cells.reshape(-1, 2)                                          # same point multiset in every cell
    rsp = g["ref_sorted_points_raw"]
    for lo, hi in flat[flat[:, 1] > flat[:, 0]][:200]:
        a = sp[lo:hi][np.lexsort(sp[lo:hi].T)]
        b = rsp[lo:hi][np.lexsort(rsp[lo:hi].T)]
        assert np.array_equal(a, b)
    assert np.array_equal(O.sort_features(feats, idx), g["ref_sort_features"])

    qp, qb, qi, qph = O.poisson_sampling(sp, sb, cells, mn, mx, rs, B, sinv, return_phase=True)
    ri, rb = g["ref_poisson_indexs_raw"], g["ref_poisson_batch_ids_raw"].reshape(-1)
    assert np.array_equal(np.sort(ri), np.sort(qi))                      # identical accepted set
    phase_of = dict(zip(qi.tolist(), qph.tolist()))
    group = np.array([int(rb[k]) * 27 + phase_of[int(ri[k])] for k in range(ri.shape[0])])
    assert np.all(np.diff(group) >= 0)                                   # reference emits (cloud, phase) groups in order
    assert np.array_equal(g["ref_poisson_points_raw"], sp[ri])

    start, pairs = O.find_neighbors(qp, qb, sp, cells, mn, mx, rg, B, sinv)
    assert np.array_equal(start, g["ref_start_indexs"])
    assert np.array_equal(pairs, g["ref_neighbors"])

    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, win, rg, B, sinv)
    assert rel_err(pdf, g["ref_pdfs"], 0.0) <= 1e-5

    w = make_conv_weights(C, Co, combin, seed=7)
    args = (sp, sf, sb, g["ref_pdfs"], qp, start, pairs, mn, mx, *w, Co, combin, B, rg, sinv, avg)
    out = O.spatial_conv(*args, accum_double=True)
    scale = O.spatial_conv(*args, abs_sum=True)
    assert out.shape == g["ref_conv_out"].shape
    assert sum_err(out, g["ref_conv_out"], scale) <= 1e-5

    assert np.array_equal(O.get_sampled_features(qi, sf), g["ref_sampled_features"])
    assert np.array_equal(O.transform_indexs(qi, idx), g["ref_transformed_indexs"])
