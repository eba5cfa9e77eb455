"""Raw, sync-free bindings of the C-ABI on torch CUDA tensors.

Each function enqueues kernels on torch's current CUDA stream and returns immediately.  Dynamic extents follow the
C-ABI convention: a tensor is allocated with a *capacity* number of rows and an optional int32[1] device tensor
(`n_dev`, `m_dev`, `k_dev`) holds the live row count.  `MCConvModule` (the reference's op API) wraps these with
exact output shapes; `MCConvBuilder`'s planned mode keeps the capacities and never blocks the host.

torch is used for device memory and streams only -- no arithmetic of the hot path happens in torch.
"""
import ctypes

import torch

from ._lib import lib, check

_i32 = torch.int32
_f32 = torch.float32


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t):
    if not t.is_cuda:
        raise RuntimeError("mccnn_b200: tensors must live on a CUDA device (there is no CPU path)")
    return t.device


def _f(t):
    if t.dtype != _f32:
        raise ValueError("expected a float32 tensor, got %s" % t.dtype)
    return t.contiguous()


def _i(t):
    if t.dtype != _i32:
        raise ValueError("expected an int32 tensor, got %s" % t.dtype)
    return t.contiguous()


def new_count(device, value=0):
    return torch.full((1,), value, dtype=_i32, device=device)


def get_block_size():
    return int(lib.mc_get_block_size())


def compute_aabb(pts, bids, batch_size, scale_inv):
    pts, bids = _f(pts), _i(bids)
    dev = _dev(pts)
    mn = torch.empty((batch_size, 3), dtype=_f32, device=dev)
    mx = torch.empty((batch_size, 3), dtype=_f32, device=dev)
    ws = torch.empty((1,), dtype=_i32, device=dev)
    check(lib.mc_compute_aabb(int(scale_inv), pts.shape[0], batch_size, _p(pts), _p(bids), _p(mn), _p(mx), _p(ws),
                              _stream()))
    return mn, mx


def num_cells(cell_size, scale_inv, aabb_min, aabb_max):
    """determineNumCells; blocks the host only when scale_inv is False (needs row 0 of the box)."""
    out = ctypes.c_int(0)
    ws = None if scale_inv else torch.empty((1,), dtype=_i32, device=aabb_min.device)
    check(lib.mc_determine_num_cells(int(scale_inv), float(cell_size), _p(aabb_min), _p(aabb_max), _p(ws),
                                     ctypes.byref(out), _stream()))
    return out.value


def verify_num_cells(cell_size, aabb_min, aabb_max, expected, flag):
    """Device-side check that the box still implies `expected` cells per axis (absolute radii); ORs 1 into `flag`."""
    check(lib.mc_verify_num_cells(float(cell_size), _p(aabb_min), _p(aabb_max), int(expected), _p(flag), _stream()))


def sort_points_step1(pts, bids, aabb_min, aabb_max, batch_size, n_cells, n_dev=None):
    pts, bids = _f(pts), _i(bids)
    dev = _dev(pts)
    N = pts.shape[0]
    keys = torch.empty((N,), dtype=_i32, device=dev)
    new_idx = torch.empty((N,), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_sort_step1_workspace_ints(N, batch_size, n_cells),), dtype=_i32, device=dev)
    check(lib.mc_sort_points_step1(N, _p(n_dev), batch_size, n_cells, _p(aabb_min), _p(aabb_max), _p(pts), _p(bids),
                                   _p(ws), _p(keys), _p(new_idx), _stream()))
    return keys, new_idx


def sort_points_step2(pts, bids, feats, keys, new_idx, batch_size, n_cells, n_dev=None):
    pts, bids, feats, keys, new_idx = _f(pts), _i(bids), _f(feats), _i(keys), _i(new_idx)
    dev = _dev(pts)
    N, C = pts.shape[0], feats.shape[1]
    out_pts = torch.empty_like(pts)
    out_bids = torch.empty_like(bids)
    out_feats = torch.empty_like(feats)
    cell_idx = torch.empty((batch_size, n_cells, n_cells, n_cells, 2), dtype=_i32, device=dev)
    ws = torch.empty((max(N, 1),), dtype=_i32, device=dev)
    check(lib.mc_sort_points_step2(N, _p(n_dev), batch_size, C, n_cells, _p(pts), _p(bids), _p(feats), _p(keys),
                                   _p(new_idx), _p(ws), _p(out_pts), _p(out_bids), _p(out_feats), _p(cell_idx),
                                   _stream()))
    return out_pts, out_bids, out_feats, cell_idx


def build_grid_supported(batch_size, n_cells):
    return bool(lib.mc_build_grid_supported(batch_size, n_cells))


def build_grid(pts, bids, feats, aabb_min, aabb_max, batch_size, n_cells, violation, n_dev=None):
    """Fused step1 + step2 for cloud-contiguous batches (one launch).  `violation` (int32[1], device) is set when the
    batch is not contiguous.  Returns (keys, new_idx, sort_pts, sort_bids, sort_feats, cell_idx)."""
    pts, bids, feats = _f(pts), _i(bids), _f(feats)
    dev = _dev(pts)
    N, C = pts.shape[0], feats.shape[1]
    keys = torch.empty((N,), dtype=_i32, device=dev)
    new_idx = torch.empty((N,), dtype=_i32, device=dev)
    out_pts = torch.empty_like(pts)
    out_bids = torch.empty_like(bids)
    out_feats = torch.empty_like(feats)
    cell_idx = torch.empty((batch_size, n_cells, n_cells, n_cells, 2), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_build_grid_workspace_ints(N),), dtype=_i32, device=dev)
    check(lib.mc_build_grid(N, _p(n_dev), batch_size, C, n_cells, _p(aabb_min), _p(aabb_max), _p(pts), _p(bids),
                            _p(feats), _p(ws), _p(violation), _p(keys), _p(new_idx), _p(out_pts), _p(out_bids),
                            _p(out_feats), _p(cell_idx), _stream()))
    return keys, new_idx, out_pts, out_bids, out_feats, cell_idx


def sort_features(feats, new_idx, n_dev=None, affine=None):
    """out[new_idx[i]] = feats[i]; with affine=(scale, shift, relu): out[new_idx[i]] = act(feats[i] * scale + shift)."""
    feats, new_idx = _f(feats), _i(new_idx)
    out = torch.empty_like(feats)
    if affine is not None:
        scale, shift, relu = affine
        check(lib.mc_sort_features_affine(feats.shape[0], _p(n_dev), feats.shape[1], _p(feats), _p(new_idx),
                                          _p(_f(scale)), _p(_f(shift)), int(bool(relu)), _p(out), _stream()))
        return out
    check(lib.mc_sort_features(feats.shape[0], _p(n_dev), feats.shape[1], _p(feats), _p(new_idx), _p(out),
                               _stream()))
    return out


def sort_features_back(feats, new_idx, n_dev=None):
    feats, new_idx = _f(feats), _i(new_idx)
    out = torch.empty_like(feats)
    check(lib.mc_sort_features_back(feats.shape[0], _p(n_dev), feats.shape[1], _p(feats), _p(new_idx), _p(out),
                                    _stream()))
    return out


def transform_indexs(idx, new_idx, m_dev=None, n_dev=None):
    idx, new_idx = _i(idx), _i(new_idx)
    dev = _dev(idx)
    out = torch.empty_like(idx)
    ws = torch.empty((max(new_idx.shape[0], 1),), dtype=_i32, device=dev)
    check(lib.mc_transform_indexs(idx.shape[0], _p(m_dev), new_idx.shape[0], _p(n_dev), _p(idx), _p(new_idx), _p(ws),
                                  _p(out), _stream()))
    return out


def find_neighbors_count(samples, sbids, pts2, cell_idx, aabb_min, aabb_max, radius, scale_inv, m_dev=None):
    """First phase: per-sample counts scanned in place; returns (start_local, workspace, k_dev)."""
    samples, sbids, pts2, cell_idx = _f(samples), _i(sbids), _f(pts2), _i(cell_idx)
    dev = _dev(samples)
    M = samples.shape[0]
    start = torch.empty((M,), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_find_neighbors_workspace_ints(M),), dtype=_i32, device=dev)
    k_dev = torch.empty((1,), dtype=_i32, device=dev)
    check(lib.mc_find_neighbors_count(int(scale_inv), M, _p(m_dev), cell_idx.shape[1], float(radius), _p(samples),
                                      _p(sbids), _p(pts2), _p(cell_idx), _p(aabb_min), _p(aabb_max), _p(start),
                                      _p(ws), _p(k_dev), _stream()))
    return start, ws, k_dev


def find_neighbors_pack(samples, sbids, pts2, cell_idx, aabb_min, aabb_max, radius, scale_inv, start, ws, capacity,
                        m_dev=None, overflow=None):
    """Second phase: finalises `start` in place and fills pairs[capacity, 2]."""
    samples, sbids, pts2, cell_idx = _f(samples), _i(sbids), _f(pts2), _i(cell_idx)
    dev = _dev(samples)
    pairs = torch.empty((capacity, 2), dtype=_i32, device=dev)
    check(lib.mc_find_neighbors_pack(int(scale_inv), samples.shape[0], _p(m_dev), cell_idx.shape[1], float(radius),
                                     _p(samples), _p(sbids), _p(pts2), _p(cell_idx), _p(aabb_min), _p(aabb_max),
                                     _p(start), _p(ws), _p(pairs), capacity, _p(overflow), _stream()))
    return pairs


def find_neighbors_fused(samples, sbids, pts2, cell_idx, aabb_min, aabb_max, radius, scale_inv, capacity, m_dev=None,
                         overflow=None):
    """Count + scan + fill in one launch.  Returns (start[M], pairs[capacity, 2], k_dev)."""
    samples, sbids, pts2, cell_idx = _f(samples), _i(sbids), _f(pts2), _i(cell_idx)
    dev = _dev(samples)
    M = samples.shape[0]
    start = torch.empty((M,), dtype=_i32, device=dev)
    pairs = torch.empty((capacity, 2), dtype=_i32, device=dev)
    k_dev = torch.empty((1,), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_find_neighbors_workspace_ints(M),), dtype=_i32, device=dev)
    check(lib.mc_find_neighbors_fused(int(scale_inv), M, _p(m_dev), cell_idx.shape[1], float(radius), _p(samples),
                                      _p(sbids), _p(pts2), _p(cell_idx), _p(aabb_min), _p(aabb_max), _p(start), _p(ws),
                                      _p(pairs), capacity, _p(k_dev), _p(overflow), _stream()))
    return start, pairs, k_dev


def find_neighbors_pdf_fused(samples, sbids, pts2, cell_idx, aabb_min, aabb_max, radius, window, scale_inv, capacity,
                             m_dev=None, overflow=None):
    """find_neighbors_fused + compute_pdf in one launch.  Returns (start[M], pairs[capacity, 2], pdf[capacity], k_dev)."""
    samples, sbids, pts2, cell_idx = _f(samples), _i(sbids), _f(pts2), _i(cell_idx)
    dev = _dev(samples)
    M = samples.shape[0]
    start = torch.empty((M,), dtype=_i32, device=dev)
    pairs = torch.empty((capacity, 2), dtype=_i32, device=dev)
    pdf = torch.empty((capacity,), dtype=_f32, device=dev)
    k_dev = torch.empty((1,), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_find_neighbors_workspace_ints(M),), dtype=_i32, device=dev)
    check(lib.mc_find_neighbors_pdf_fused(int(scale_inv), float(window), M, _p(m_dev), cell_idx.shape[1], float(radius),
                                          _p(samples), _p(sbids), _p(pts2), _p(cell_idx), _p(aabb_min), _p(aabb_max),
                                          _p(start), _p(ws), _p(pairs), _p(pdf), capacity, _p(k_dev), _p(overflow),
                                          _stream()))
    return start, pairs, pdf, k_dev


def compute_pdf(pts, bids, aabb_min, aabb_max, start, pairs, window, radius, scale_inv, m_dev=None, k_dev=None):
    pts, bids, start, pairs = _f(pts), _i(bids), _i(start), _i(pairs)
    dev = _dev(pts)
    K = pairs.shape[0]
    pdf = torch.empty((K,), dtype=_f32, device=dev)
    check(lib.mc_compute_pdf(int(scale_inv), float(window), start.shape[0], _p(m_dev), K, _p(k_dev), float(radius),
                             _p(pts), _p(bids), _p(aabb_min), _p(aabb_max), _p(start), _p(pairs), _p(pdf),
                             _stream()))
    return pdf


def poisson_sampling(pts, bids, cell_idx, aabb_min, aabb_max, radius, batch_size, scale_inv, n_dev=None,
                     capacity=None, overflow=None):
    """Returns (out_pts[cap,3], out_bids[cap], out_idx[cap], m_dev)."""
    pts, bids, cell_idx = _f(pts), _i(bids), _i(cell_idx)
    dev = _dev(pts)
    N = pts.shape[0]
    cap = N if capacity is None else min(capacity, N)
    out_pts = torch.empty((cap, 3), dtype=_f32, device=dev)
    out_bids = torch.empty((cap,), dtype=_i32, device=dev)
    out_idx = torch.empty((cap,), dtype=_i32, device=dev)
    m_dev = torch.empty((1,), dtype=_i32, device=dev)
    ws = torch.empty((lib.mc_poisson_workspace_ints(N, batch_size),), dtype=_i32, device=dev)
    check(lib.mc_poisson_sampling(int(scale_inv), float(radius), N, _p(n_dev), batch_size, cell_idx.shape[1],
                                  _p(aabb_min), _p(aabb_max), _p(pts), _p(bids), _p(cell_idx), _p(ws), cap,
                                  _p(out_pts), _p(out_bids), _p(out_idx), _p(m_dev), _p(overflow), _stream()))
    return out_pts, out_bids, out_idx, m_dev


def get_sampled_features(idx, feats, m_dev=None):
    idx, feats = _i(idx), _f(feats)
    out = torch.empty((idx.shape[0], feats.shape[1]), dtype=_f32, device=_dev(feats))
    check(lib.mc_get_sampled_features(idx.shape[0], _p(m_dev), feats.shape[1], _p(idx), _p(feats), _p(out),
                                      _stream()))
    return out


def spatial_conv(pts, feats, bids, pdfs, samples, start, pairs, aabb_min, aabb_max, w1, b1, w2, b2, wout, bout,
                 num_out_features, combin, radius, scale_inv, avg, m_dev=None, k_dev=None):
    """Parameters in the OP's order (W1, b1, W2, b2, Wout, bout), flat reference layouts."""
    pts, feats, bids, pdfs, samples = _f(pts), _f(feats), _i(bids), _f(pdfs), _f(samples)
    start, pairs = _i(start), _i(pairs)
    w1, b1, w2, b2, wout, bout = (_f(t) for t in (w1, b1, w2, b2, wout, bout))
    dev = _dev(pts)
    M, K, C = samples.shape[0], pairs.shape[0], feats.shape[1]
    out = torch.empty((M, num_out_features), dtype=_f32, device=dev)
    ws_n = lib.mc_spatial_conv_workspace_floats(M, C, num_out_features, int(combin))
    ws = torch.empty((ws_n,), dtype=_f32, device=dev) if ws_n > 0 else None
    check(lib.mc_spatial_conv(int(avg), int(scale_inv), K, _p(k_dev), C, num_out_features, M, _p(m_dev), int(combin),
                              float(radius), aabb_min.shape[0], _p(pts), _p(bids), _p(feats), _p(pdfs), _p(samples), _p(start),
                              _p(pairs), _p(aabb_min), _p(aabb_max), _p(w1), _p(b1), _p(w2), _p(b2), _p(wout),
                              _p(bout), _p(ws), _p(out), _stream()))
    return out


def spatial_conv_grad(pts, feats, bids, pdfs, samples, start, pairs, aabb_min, aabb_max, w1, b1, w2, b2, wout, bout,
                      out_grad, num_out_features, combin, radius, scale_inv, avg):
    """-> (featureGrads [N, C], dW1, db1, dW2, db2, dWout, dbout), each shaped like its parameter."""
    pts, feats, bids, pdfs, samples = _f(pts), _f(feats), _i(bids), _f(pdfs), _f(samples)
    start, pairs, out_grad = _i(start), _i(pairs), _f(out_grad)
    w1, b1, w2, b2, wout, bout = (_f(t) for t in (w1, b1, w2, b2, wout, bout))
    dev = _dev(pts)
    M, K, C, N = samples.shape[0], pairs.shape[0], feats.shape[1], pts.shape[0]
    d_feat = torch.empty((N, C), dtype=_f32, device=dev)
    grads = [torch.empty_like(t) for t in (w1, b1, w2, b2, wout, bout)]
    ws = torch.empty((lib.mc_spatial_conv_grad_workspace_floats(K, C, num_out_features, int(combin)),), dtype=_f32,
                     device=dev)
    check(lib.mc_spatial_conv_grad(int(avg), int(scale_inv), K, C, num_out_features, M, N, int(combin), float(radius),
                                   aabb_min.shape[0], _p(pts), _p(bids), _p(feats), _p(pdfs), _p(samples), _p(start),
                                   _p(pairs), _p(aabb_min), _p(aabb_max), _p(w1), _p(b1), _p(w2), _p(b2), _p(wout),
                                   _p(bout), _p(out_grad), _p(ws), _p(d_feat), _p(grads[0]), _p(grads[1]),
                                   _p(grads[2]), _p(grads[3]), _p(grads[4]), _p(grads[5]), _stream()))
    return (d_feat,) + tuple(grads)


def sort_points_step2_grad(new_idx, out_pts_grad, out_feat_grad):
    new_idx, gp, gf = _i(new_idx), _f(out_pts_grad), _f(out_feat_grad)
    ip, i_f = torch.empty_like(gp), torch.empty_like(gf)
    check(lib.mc_sort_points_step2_grad(new_idx.shape[0], gf.shape[1], _p(gp), _p(gf), _p(new_idx), _p(ip), _p(i_f),
                                        _stream()))
    return ip, i_f


def get_sampled_features_grad(idx, feats, out_grad):
    idx, feats, out_grad = _i(idx), _f(feats), _f(out_grad)
    g = torch.empty_like(feats)
    check(lib.mc_get_sampled_features_grad(feats.shape[0], feats.shape[1], idx.shape[0], _p(idx), _p(out_grad), _p(g),
                                           _stream()))
    return g


def scale_shift_act(x, scale, shift, relu, n_dev=None):
    """y = act(x * scale + shift) per column (inference batch norm + ReLU), live rows only."""
    x, scale, shift = _f(x), _f(scale), _f(shift)
    y = torch.empty_like(x)
    check(lib.mc_scale_shift_act(x.shape[0], _p(n_dev), x.shape[1], _p(x), _p(scale), _p(shift), int(relu), _p(y),
                                 _stream()))
    return y


def dense_layer(x, w, bias=None, pre=None, pre_relu=False, post=None, post_relu=False, n_dev=None, rows_hint=0):
    """y = act2((act1(x * pre[0] + pre[1]) @ w + bias) * post[0] + post[1]); w [K, N] row-major."""
    x, w = _f(x), _f(w)
    K, N = w.shape
    if x.shape[1] != K:
        raise ValueError("dense_layer: input has %d columns, weights expect %d" % (x.shape[1], K))
    y = torch.empty((x.shape[0], N), dtype=_f32, device=_dev(x))
    s1, t1 = (None, None) if pre is None else (_f(pre[0]), _f(pre[1]))
    s2, t2 = (None, None) if post is None else (_f(post[0]), _f(post[1]))
    b = None if bias is None else _f(bias)
    check(lib.mc_dense_layer(x.shape[0], _p(n_dev), int(rows_hint or 0), K, N, _p(x), _p(s1), _p(t1), int(pre_relu), _p(w), _p(b), _p(s2),
                             _p(t2), int(post_relu), _p(y), _stream()))
    return y


def dense_layer_fewrows(x, w, bias=None, pre=None, pre_relu=False, post=None, post_relu=False, n_dev=None):
    """dense_layer for few live rows (<= ~128: one per cloud) and layers of any width: W is read once for all rows."""
    x, w = _f(x), _f(w)
    K, N = w.shape
    if x.shape[1] != K:
        raise ValueError("dense_layer_fewrows: input has %d columns, weights expect %d" % (x.shape[1], K))
    y = torch.empty((x.shape[0], N), dtype=_f32, device=_dev(x))
    nws = lib.mc_dense_fewrows_workspace_floats(x.shape[0], K, N)
    ws = torch.empty((nws,), dtype=_f32, device=_dev(x)) if nws else None
    s1, t1 = (None, None) if pre is None else (_f(pre[0]), _f(pre[1]))
    s2, t2 = (None, None) if post is None else (_f(post[0]), _f(post[1]))
    b = None if bias is None else _f(bias)
    check(lib.mc_dense_layer_fewrows(x.shape[0], _p(n_dev), K, N, _p(x), _p(s1), _p(t1), int(pre_relu), _p(w), _p(b),
                                     _p(s2), _p(t2), int(post_relu), _p(ws), _p(y), _stream()))
    return y


def split_tf32_transposed(w):
    """(hi, lo) halves of w TRANSPOSED to [N, K] for dense_layer_tc: hi = tf32(w^T), lo = w^T - hi."""
    wt = _f(w).t().contiguous()
    hi, lo = torch.empty_like(wt), torch.empty_like(wt)
    check(lib.mc_split_tf32(wt.numel(), _p(wt), _p(hi), _p(lo), _stream()))
    return hi, lo


def dense_layer_tc_supported(k, n):
    return bool(lib.mc_dense_layer_tc_supported(int(k), int(n)))


def dense_layer_tc(x, wt_hi, wt_lo, bias=None, pre=None, pre_relu=False, post=None, post_relu=False, n_dev=None):
    """dense_layer on the tcgen05 tensor cores; wt_hi / wt_lo [N, K] from split_tf32_transposed(w)."""
    x, wt_hi, wt_lo = _f(x), _f(wt_hi), _f(wt_lo)
    N, K = wt_hi.shape
    if x.shape[1] != K:
        raise ValueError("dense_layer_tc: input has %d columns, weights expect %d" % (x.shape[1], K))
    y = torch.empty((x.shape[0], N), dtype=_f32, device=_dev(x))
    s1, t1 = (None, None) if pre is None else (_f(pre[0]), _f(pre[1]))
    s2, t2 = (None, None) if post is None else (_f(post[0]), _f(post[1]))
    b = None if bias is None else _f(bias)
    check(lib.mc_dense_layer_tc(x.shape[0], _p(n_dev), K, N, _p(x), _p(s1), _p(t1), int(pre_relu), _p(wt_hi), _p(wt_lo),
                                _p(b), _p(s2), _p(t2), int(post_relu), _p(y), _stream()))
    return y


def mlp_rows(x, layers, pre=None, pre_relu=False, n_dev=None):
    """A whole MLP on few rows in one launch: `layers` = [(w [K, N], bias | None, post (scale, shift) | None,
    post_relu), ...]; a_0 = act(x * pre[0] + pre[1]).  Every width must be <= 1024."""
    import ctypes
    x = _f(x)
    n = len(layers)
    dims = [x.shape[1]] + [int(l[0].shape[1]) for l in layers]
    keep = []

    def ptr(t):
        if t is None:
            return None
        t = _f(t)
        keep.append(t)
        return t.data_ptr()

    W = (ctypes.c_void_p * n)(*[ptr(l[0]) for l in layers])
    B = (ctypes.c_void_p * n)(*[ptr(l[1]) for l in layers])
    S = (ctypes.c_void_p * n)(*[ptr(l[2][0]) if l[2] is not None else None for l in layers])
    T = (ctypes.c_void_p * n)(*[ptr(l[2][1]) if l[2] is not None else None for l in layers])
    R = (ctypes.c_int * n)(*[int(bool(l[3])) for l in layers])
    D = (ctypes.c_int * (n + 1))(*dims)
    for l, (k_in, layer) in enumerate(zip(dims[:-1], layers)):
        if layer[0].shape[0] != k_in:
            raise ValueError("mlp_rows: layer %d expects %d inputs, got %d" % (l, layer[0].shape[0], k_in))
    y = torch.empty((x.shape[0], dims[-1]), dtype=_f32, device=_dev(x))
    s0, t0 = (None, None) if pre is None else (_f(pre[0]), _f(pre[1]))
    check(lib.mc_mlp_rows(x.shape[0], _p(n_dev), n, D, _p(x), _p(s0), _p(t0), int(bool(pre_relu)), W, B, S, T, R,
                          _p(y), _stream()))
    return y


def allgather_rows_p2p(local, peer_ptrs, counter, rank, world, out):
    """All-gather of `local` (same numel on every rank) through symmetric peer buffers (csrc/peer_gather.cu)."""
    local = _f(local)
    check(lib.mc_allgather_rows_p2p(local.numel(), int(rank), int(world), _p(local), _p(peer_ptrs), _p(counter),
                                    _p(out), _stream()))
    return out


def peer_put(local, peer_ptrs, state, rank, world, slots, cap_rows, n_dev=None, flag=None):
    """Step of the put / collect exchange (csrc/peer_gather.cu): this rank's live rows -> every rank's symmetric buffer."""
    local = _f(local)
    check(lib.mc_peer_put(int(cap_rows), _p(n_dev), local.shape[1], int(rank), int(world), int(slots), _p(local),
                          _p(flag), _p(peer_ptrs), _p(state), _stream()))


def peer_collect(peer_ptrs, state, rank, world, slots, cap_rows, cols, out, meta, compact, pending_only=False):
    """pending_only: do nothing unless a put of this exchange is still uncollected (decided on the device)."""
    check(lib.mc_peer_collect(int(cap_rows), int(cols), int(rank), int(world), int(slots),
                              int(bool(compact)) | (2 if pending_only else 0),
                              _p(peer_ptrs), _p(state), _p(out), _p(meta), _stream()))
    return out, meta
