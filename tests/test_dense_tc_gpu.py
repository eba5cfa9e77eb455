"""mc_dense_layer_tc (csrc/dense_tc.cu: tcgen05 + TMEM + TMA, FP32 through the 3xTF32 split) against the float64 torch
formulation of the reference's TensorFlow glue (utils/MCNetworkUtils.py:109-142: batch norm, ReLU, matmul + bias),
1e-5 relative to the sum of |terms| of every output, on the shapes of the classification and segmentation networks
(MCSeg's 131 072-row 1x1 layers included), ragged K / N tails and device-side live row counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
SHAPES = [(5400, 64, 128), (387, 128, 256), (130, 32, 16), (1000, 272, 256), (4096, 36, 50), (999, 100, 40),
          (2000, 1664, 512), (131072, 832, 256), (131072, 256, 50), (100000, 128, 21), (20000, 1536, 512)]


def _case(torch, dev, rows, K, N, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    x = r(rows, K).to(dev)
    w = (r(K, N) / np.sqrt(K)).to(dev)
    b = (0.1 * r(N)).to(dev)
    s1, t1 = (0.5 + torch.rand(K, generator=g)).to(dev), (0.1 * r(K)).to(dev)
    s2, t2 = (0.5 + torch.rand(N, generator=g)).to(dev), (0.1 * r(N)).to(dev)
    return x, w, b, s1, t1, s2, t2


@pytest.mark.parametrize("rows,K,N", SHAPES)
def test_dense_tc_matches_float64(rows, K, N):
    import torch
    from mccnn_b200 import ops
    dev = torch.device("cuda:0")
    x, w, b, s1, t1, s2, t2 = _case(torch, dev, rows, K, N, seed=rows + K + N)
    hi, lo = ops.split_tf32_transposed(w)
    assert torch.equal((hi + lo).t(), w) and int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    got = ops.dense_layer_tc(x, hi, lo, b, pre=(s1, t1), pre_relu=True, post=(s2, t2), post_relu=True)
    plain = ops.dense_layer_tc(x, hi, lo)
    torch.cuda.synchronize()
    worst = 0.0
    for r0 in range(0, rows, 16384):   # float64 reference in row chunks
        xs = x[r0:r0 + 16384].double()
        a = torch.relu(xs * s1.double() + t1.double())
        want = torch.relu((a @ w.double() + b.double()) * s2.double() + t2.double())
        scale = (a.abs() @ w.double().abs() + b.double().abs()) * s2.double() + t2.double().abs()
        worst = max(worst, float(((got[r0:r0 + 16384].double() - want).abs() / scale).max()))
        ref = xs @ w.double()
        e2 = float(((plain[r0:r0 + 16384].double() - ref).abs() / (xs.abs() @ w.double().abs() + 1e-30)).max())
        worst = max(worst, e2)
    print("dense_tc %d x %d -> %d: max error / sum|terms| = %.2e" % (rows, K, N, worst))
    assert worst <= RTOL, worst


def test_dense_tc_live_rows_and_determinism():
    import torch
    from mccnn_b200 import ops
    dev = torch.device("cuda:0")
    x, w, b, s1, t1, s2, t2 = _case(torch, dev, 3000, 128, 96, seed=5)
    hi, lo = ops.split_tf32_transposed(w)
    full = ops.dense_layer_tc(x, hi, lo, b, post=(s2, t2), post_relu=True)
    again = ops.dense_layer_tc(x, hi, lo, b, post=(s2, t2), post_relu=True)
    assert torch.equal(full, again)
    for live in (1, 127, 128, 129, 2999):
        cnt = torch.tensor([live], dtype=torch.int32, device=dev)
        part = ops.dense_layer_tc(x, hi, lo, b, post=(s2, t2), post_relu=True, n_dev=cnt)
        assert torch.equal(part[:live], full[:live]), live
    assert ops.dense_layer_tc_supported(64, 7) and not ops.dense_layer_tc_supported(30, 8)
    assert not ops.dense_layer_tc_supported(16, 8)


def test_dense_tc_timing_vs_library_fp32():
    """Not a pass/fail bar: prints the kernel's time next to torch.addmm (cuBLAS FP32, what the route used before)."""
    import torch
    from mccnn_b200 import ops
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    for rows, K, N in ((131072, 832, 256), (131072, 1536, 512), (131072, 256, 128), (640000, 128, 128), (5400, 64, 128)):
        x, w, b, s1, t1, s2, t2 = _case(torch, dev, rows, K, N, seed=1)
        hi, lo = ops.split_tf32_transposed(w)
        def tc():
            return ops.dense_layer_tc(x, hi, lo, b, pre=(s1, t1), pre_relu=True)
        def lib():
            return torch.addmm(b, ops.scale_shift_act(x, s1, t1, True), w)
        res = {}
        for name, fn in (("tc", tc), ("cublas", lib)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        tf = 2.0 * rows * K * N / (res["tc"] * 1e-3) / 1e12
        print("dense %7d x %4d -> %3d: tcgen05 3xTF32 %.3f ms (%.1f TFLOP/s fp32-equivalent), BN+ReLU launch + cuBLAS "
              "FP32 %.3f ms" % (rows, K, N, res["tc"], tf, res["cublas"]))
