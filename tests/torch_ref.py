"""float64 torch restatement of spatial_conv (tf_ops/spatial_conv.cu:49-78,131-175 / :201-229,280-324) that runs on
the GPU, for parity checks at sizes where the single-threaded C oracle would take minutes.  TEST INFRASTRUCTURE.

Returns the order-free float64 value of every output AND the error scale of the float32 sum behind it: the sum of
|contributions| with magnitudes propagated through the three MLP layers -- exactly `accumDouble = 1` and `= 2` of
oracle/mccnn_oracle.c::orc_spatial_conv, against which tests/test_fullsize_gpu.py pins this function on small cases
before using it on the large ones.  Pairs are processed in chunks so that K x neurons intermediates stay bounded."""
import torch


def spatial_conv_fp64(pts, feats, bids, pdfs, samples, start, pairs, mn, mx, w1, w2, wo, b1, b2, bo, num_out, combin,
                      radius, scale_inv, avg, chunk=1 << 18):
    """Arguments as the reference's python spatial_conv (MCConvModuleSrc:66-73) minus batchSize; CUDA or CPU tensors.
    -> (out float64 [M, Co], scale float64 [M, Co])."""
    dev = pts.device
    M, K, C = samples.shape[0], pairs.shape[0], feats.shape[1]
    n_out = C * num_out if combin else C
    nb = (n_out + 7) // 8
    out_c = num_out if combin else C
    f64 = torch.float64
    W1 = w1.reshape(-1)[: nb * 24].reshape(nb, 8, 3).to(f64)          # [(8q+t)*3 + a]
    B1 = b1.reshape(-1)[: nb * 8].reshape(nb, 8).to(f64)
    W2 = w2.reshape(-1)[: nb * 64].reshape(nb, 8, 8).to(f64)          # [q*64 + t*8 + u]
    B2 = b2.reshape(-1)[: nb * 8].reshape(nb, 8).to(f64)
    WO = wo.reshape(-1)[: nb * 64].reshape(nb, 8, 8).to(f64)          # [(8q+t)*8 + u]
    BO = bo.reshape(-1)[: nb * 8].reshape(nb, 8).to(f64)
    ext = (mx - mn).max(dim=1).values                                # float32, as the kernels
    rb = radius * ext if scale_inv else torch.full_like(ext, radius)
    st = start.reshape(-1).long()
    end = torch.cat([st[1:], torch.tensor([K], device=dev, dtype=torch.long)])
    cnt = (end - st).to(torch.float32) if avg else torch.ones((M,), dtype=torch.float32, device=dev)
    out = torch.zeros((M, out_c), dtype=f64, device=dev)
    scale = torch.zeros((M, out_c), dtype=f64, device=dev)
    neuron = torch.arange(nb * 8, device=dev)
    valid = neuron < n_out
    col_c = (neuron % C).clamp(max=C - 1)
    col_o = torch.div(neuron, C, rounding_mode="floor").clamp(max=out_c - 1) if combin else col_c
    bflat = bids.reshape(-1).long()
    pdf = pdfs.reshape(-1)
    for p0 in range(0, K, chunk):
        pr = pairs[p0:p0 + chunk].long()
        j, i = pr[:, 0], pr[:, 1]
        r = rb[bflat[j]]
        d = ((pts[j] - samples[i]) / r[:, None]).to(f64)             # float32 division like the kernel, then widened
        w = (pdf[p0:p0 + chunk] * cnt[i]).to(f64)                   # float32 product pdf * cnt, as the kernel
        pre1 = torch.einsum("pa,qta->pqt", d, W1) + B1
        h1 = pre1.clamp(min=0)
        m1 = torch.einsum("pa,qta->pqt", d.abs(), W1.abs()) + B1.abs()
        h2 = (torch.einsum("pqu,qtu->pqt", h1, W2) + B2).clamp(min=0)
        m2 = torch.einsum("pqu,qtu->pqt", m1, W2.abs()) + B2.abs()
        kap = (torch.einsum("pqu,qtu->pqt", h2, WO) + BO).reshape(-1, nb * 8)
        mk = (torch.einsum("pqu,qtu->pqt", m2, WO.abs()) + BO.abs()).reshape(-1, nb * 8)
        f = feats[j][:, col_c].to(f64)
        contrib = torch.where(valid, f * kap / w[:, None], torch.zeros((), dtype=f64, device=dev))
        mag = torch.where(valid, f.abs() * mk / w.abs()[:, None], torch.zeros((), dtype=f64, device=dev))
        if combin and C > 1:  # several neurons feed one output: sum them per pair first
            cc = torch.zeros((contrib.shape[0], out_c), dtype=f64, device=dev).index_add_(1, col_o, contrib)
            mm = torch.zeros((contrib.shape[0], out_c), dtype=f64, device=dev).index_add_(1, col_o, mag)
        else:
            cc, mm = contrib[:, :out_c], mag[:, :out_c]
        out.index_add_(0, i, cc)
        scale.index_add_(0, i, mm)
    return out, scale
