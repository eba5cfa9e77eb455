"""Generates tests/golden/*.npz: outputs of the REFERENCE's own CUDA kernels (oracle/_ref/libmccnn_ref.so, built by
oracle/Makefile from /root/reference/tf_ops/*.cu, unmodified) on small seeded inputs.  Run on a GPU box:

    python tests/golden/make_golden.py            # writes gpurun_out/golden/*.npz (copy them to tests/golden/)

The reference has no tests or golden vectors of its own; these files pin the CPU oracle
(tests/test_oracle_golden.py, runs without a GPU) to what the reference kernels actually compute on a B200.

Inputs to the order-sensitive ops (find_neighbors, compute_pdf, poisson_sampling, spatial_conv) are the CANONICALLY
sorted arrays (stable counting sort, produced by the CPU oracle) so that their reference outputs are deterministic;
the reference's own sort output (an arbitrary within-cell permutation, tf_ops/sort_gpu.cu:170) is stored raw.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from common import make_batch, make_conv_weights  # noqa: E402
from oracle import cpu as O  # noqa: E402
from oracle import ref as R  # noqa: E402

CASES = {
    # name: (batch, points, dist, ragged, grid radius, sample radius, window, C, Co, combin, scaleInv, avg)
    "cube_b2_n256": (2, 256, "cube", False, 0.2, 0.1, 0.2, 3, 8, True, True, True),
    "sphere_b3_ragged": (3, 300, "sphere", True, 0.25, 0.12, 0.25, 8, 8, False, True, True),
    "config1_sphere_1024": (1, 1024, "sphere", False, 0.1, 0.1, 0.2, 3, 8, True, True, True),
    "room_abs_b2": (2, 1024, "room", False, 0.5, 0.3, 0.25, 1, 16, True, False, False),
}


def main():
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    A = lambda t: t.detach().cpu().numpy()  # noqa: E731
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (B, N, dist, ragged, rg, rs, win, C, Co, combin, sinv, avg) in CASES.items():
        pts, bids, feats = make_batch(B, N, dist, seed=42, num_features=C, ragged=ragged)
        if pts.shape[0] % 256 != 0 and pts.shape[0] % 256 < B:  # dodge the reference's aabb smem bug (aabb_gpu.cu:69-78)
            pts, bids, feats = pts[:-(pts.shape[0] % 256)], bids[:-(pts.shape[0] % 256)], feats[:-(pts.shape[0] % 256)]
        g = {"points": pts, "batch_ids": bids, "features": feats,
             "params": np.array([B, rg, rs, win, C, Co, int(combin), int(sinv), int(avg)], np.float64)}
        tp, tb, tf = T(pts), T(bids), T(feats)
        mn, mx = R.compute_aabb(tp, tb, B, sinv)
        g["aabb_min"], g["aabb_max"] = A(mn), A(mx)
        keys, idx = R.sort_points_step1(tp, tb, mn, mx, B, rg, sinv)
        sp, sb, sf, cells = R.sort_points_step2(tp, tb, tf, keys, idx, mn, mx, B, rg, sinv)
        g["ref_keys"], g["ref_indexs_raw"], g["ref_cell_indexs"] = A(keys), A(idx), A(cells)
        g["ref_sorted_points_raw"], g["ref_sorted_features_raw"] = A(sp), A(sf)
        # canonical sorted arrays -> deterministic reference outputs downstream
        ok, oi = O.sort_points_step1(pts, bids, A(mn), A(mx), B, rg, sinv)
        csp, csb, csf, ccells = O.sort_points_step2(pts, bids, feats, ok, oi, A(mn), A(mx), B, rg, sinv)
        assert np.array_equal(ccells, A(cells))
        tsp, tsb, tsf, tcells = T(csp), T(csb), T(csf), T(ccells)
        qp, qb, qi = R.poisson_sampling(tsp, tsb, tcells, mn, mx, rs, B, sinv)
        g["ref_poisson_indexs_raw"], g["ref_poisson_batch_ids_raw"], g["ref_poisson_points_raw"] = A(qi), A(qb), A(qp)
        # canonical sample order for the ops below
        cqp, cqb, cqi = O.poisson_sampling(csp, csb, ccells, A(mn), A(mx), rs, B, sinv)
        assert np.array_equal(np.sort(cqi), np.sort(A(qi)))
        tqp, tqb = T(cqp), T(cqb)
        start, pairs = R.find_neighbors(tqp, tqb, tsp, tcells, mn, mx, rg, B, sinv)
        g["ref_start_indexs"], g["ref_neighbors"] = A(start), A(pairs)
        pdf = R.compute_pdf(tsp, tsb, mn, mx, start, pairs, win, rg, B, sinv)
        g["ref_pdfs"] = A(pdf)
        w = make_conv_weights(C, Co, combin, seed=7)
        conv = R.spatial_conv(tsp, tsf, tsb, pdf, tqp, start, pairs, mn, mx, *[T(a) for a in w], Co, combin, B, rg,
                              sinv, avg)
        g["ref_conv_out"] = A(conv)
        g["ref_sampled_features"] = A(R.get_sampled_features(T(cqi), tsf))
        g["ref_transformed_indexs"] = A(R.transform_indexs(T(cqi), T(oi)))
        g["ref_sort_features"] = A(R.sort_features(tf, T(oi)))
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **g)
        print(name, "N", pts.shape[0], "M", cqi.shape[0], "K", A(pairs).shape[0])


if __name__ == "__main__":
    main()
