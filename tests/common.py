"""Seeded synthetic inputs shared by the parity tests, golden-vector script, smoke() and bench.py
(uniform-in-cube and sphere-surface clouds, Glorot weights, biases U(-0.1, 0.1))."""
import math

import numpy as np


def make_cloud(n, dist, rng):
    if dist == "cube":
        return rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
    if dist == "sphere":  # ModelNet-like: points on a surface
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return (0.5 * v).astype(np.float32)
    if dist == "room":  # ScanNet-like: faces of a 6 x 4 x 2.5 m box, density thinned along x (utils/DataSet.py:464-466)
        L = np.array([6.0, 4.0, 2.5])
        out = []
        while sum(len(o) for o in out) < n:
            m = 4 * n
            face = rng.integers(0, 6, m)
            p = rng.uniform(-0.5, 0.5, (m, 3)) * L
            ax = face // 2
            p[np.arange(m), ax] = np.where(face % 2 == 0, -0.5, 0.5) * L[ax]
            keep = rng.uniform(size=m) < np.sqrt(np.clip((p[:, 0] / L[0] + 0.5 - 0.2) / 0.6, 0.01, 1.0))
            out.append(p[keep])
        return np.concatenate(out)[:n].astype(np.float32)
    if dist == "lattice":  # quantised coordinates (multiples of 1/8 in [-1, 1]): duplicates, points exactly on cell
        # boundaries and pair distances exactly equal to common radii -- the float predicates' boundary cases
        return (rng.integers(0, 17, (n, 3)).astype(np.float32) / np.float32(8.0) - np.float32(1.0)).astype(np.float32)
    raise ValueError(dist)


def make_batch(batch, n, dist="sphere", seed=0, num_features=1, ragged=False):
    """Packed batch: points [sum N, 3], batch ids int32 [sum N, 1] (cloud-contiguous, utils/DataSet.py:816-818),
    features [sum N, C] (ones when C == 1, as utils/DataSet.py:822, else U(-1, 1))."""
    pts, bids = [], []
    for b in range(batch):
        rng = np.random.default_rng(seed * 1000 + b)
        nb = n if not ragged else max(1, int(n * rng.uniform(0.3, 1.0)))
        pts.append(make_cloud(nb, dist, rng))
        bids.append(np.full((nb, 1), b, np.int32))
    pts = np.concatenate(pts)
    bids = np.concatenate(bids)
    rng = np.random.default_rng(seed + 77)
    if num_features == 1:
        feats = np.ones((pts.shape[0], 1), np.float32)
    else:
        feats = rng.uniform(-1, 1, (pts.shape[0], num_features)).astype(np.float32)
    return pts, bids, feats


def make_conv_weights(num_in, num_out, combin, seed=0, block=8):
    """TF-shaped parameter tensors of one convolution (utils/MCConvBuilder.py:397-419), python-API order."""
    rng = np.random.default_rng(seed + 1234)
    n_out = num_in * num_out if combin else num_in
    nb = (n_out + block - 1) // block

    def glorot(shape, fi, fo):
        lim = math.sqrt(3.0 / ((fi + fo) / 2.0))
        return rng.uniform(-lim, lim, shape).astype(np.float32)

    w1 = glorot((3, block * nb), 3, block * nb)
    w2 = glorot((nb, block, block), block * nb, block * nb).reshape(block, nb * block)
    w3 = glorot((nb, block, block), block * nb, block * nb).reshape(block, nb * block)
    b1 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    b2 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    b3 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    return w1, w2, w3, b1, b2, b3


def rel_err(a, b, atol_scale=1e-6):
    """max |a-b| / (max(|a|,|b|) + atol), atol = atol_scale * max|b| (guards exact cancellations)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    atol = atol_scale * (np.abs(b).max() if b.size else 1.0)
    return float((np.abs(a - b) / (np.maximum(np.abs(a), np.abs(b)) + atol)).max())


def sum_err(a, b, scale):
    """max |a-b| / scale, scale = sum of |terms| behind each element: the relative error of a float32 SUM.
    (Outputs of spatial_conv are signed sums that cancel; see oracle/mccnn_oracle.c orc_spatial_conv.)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    s = np.asarray(scale, np.float64)
    if a.size == 0:
        return 0.0
    return float((np.abs(a - b) / (s + 1e-30 + 1e-12 * s.max())).max())


def named_variable(name, shape, seed=0):
    """Deterministic value of the model variable `name` (float32 ndarray): weights Glorot-uniform with TensorFlow's
    fan rule (variance_scaling factor 1, FAN_AVG, uniform -- utils/MCConvBuilder.py:394), everything that the reference
    initialises to a constant (biases 0, gamma 1, beta 0, moving mean 0 / variance 1) is drawn non-trivially so that
    every parameter path is exercised.  Seeded by the NAME, so the reference-model runner (oracle/ref_models.py) and the
    GPU tests (tests/test_models_gpu.py) obtain identical parameters without storing them."""
    import zlib
    rng = np.random.default_rng([zlib.crc32(name.encode()), seed])
    shape = tuple(int(s) for s in shape)
    if name.endswith("/gamma") or name.endswith("/moving_variance"):
        return rng.uniform(0.75, 1.25, shape).astype(np.float32)
    if name.endswith("/beta") or name.endswith("/moving_mean") or "biases" in name:
        return rng.uniform(-0.1, 0.1, shape).astype(np.float32)
    if len(shape) == 1:
        fi = fo = float(shape[0])
    elif len(shape) == 2:
        fi, fo = float(shape[0]), float(shape[1])
    else:
        rf = float(np.prod(shape[:-2]))
        fi, fo = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(3.0 / ((fi + fo) / 2.0))
    return rng.uniform(-lim, lim, shape).astype(np.float32)


# ------------------------------------------------------------------------------------------------ model cases
# Small configurations of every network of the reference (models/*.py) shared by the golden-vector generator
# (tests/golden/make_model_golden.py, runs the reference's own model files) and the GPU tests.
MODEL_CASES = {
    #  name          : (batch, points per cloud, distribution, input features, k, extra)
    "MCClassS": dict(B=3, N=512, dist="sphere", cin=1, k=8, ragged=True, ncat=10),
    "MCClass": dict(B=2, N=768, dist="sphere", cin=1, k=8, ragged=False, ncat=10),
    "MCClassH": dict(B=2, N=768, dist="cube", cin=1, k=8, ragged=True, ncat=10),
    "MCSeg": dict(B=2, N=1024, dist="sphere", cin=1, k=8, ragged=False, ncat=4, nparts=6),
    "MCSegScanNet": dict(B=2, N=6000, dist="room", cin=4, k=8, ragged=True, nsem=5),
    "MCNorm": dict(B=2, N=768, dist="sphere", cin=3, k=8, ragged=False),
    "MCNormS": dict(B=2, N=512, dist="cube", cin=3, k=8, ragged=True),
}


def model_case(name):
    """-> (inputs dict of numpy arrays, positional args AFTER (points, batchIds, features) of create_network)."""
    c = MODEL_CASES[name]
    pts, bids, feats = make_batch(c["B"], c["N"], c["dist"], seed=41 + len(name), num_features=c["cin"],
                                  ragged=c["ragged"])
    inputs = dict(points=pts, batchIds=bids, features=feats)
    B, k, cin = c["B"], c["k"], c["cin"]
    if name in ("MCClassS", "MCClass", "MCClassH"):
        args = (cin, B, k, c["ncat"], False, 1.0, 1.0)
    elif name == "MCSeg":
        inputs["catLabels"] = (bids % c["ncat"]).astype(np.int32)  # one object category per cloud, per point
        args = (cin, c["ncat"], c["nparts"], B, k, False, 1.0, 1.0)  # after catLabels
    elif name == "MCSegScanNet":
        args = (cin, c["nsem"], B, k, False, 1.0, 1.0)
    else:
        args = (cin, B, k, False)
    return inputs, args


# ------------------------------------------------------------------------------------------------ data-set cases
def make_dataset_models(num=10, seed=0):
    """Synthetic in-memory models for the DataSet tests: (points f64 [n,3], normals f64 [n,3], features f64 [n,4]
    = normal + one scalar, labels i32 [n,1], category)."""
    models = []
    for m in range(num):
        rng = np.random.default_rng([seed, m])
        n = int(rng.integers(300, 900))
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * rng.uniform(0.3, 1.0, (1, 3))            # an ellipsoid: a distinct largest axis per model
        feats = np.concatenate([v, rng.uniform(-1, 1, (n, 1))], axis=1)
        labels = rng.integers(0, 5, (n, 1)).astype(np.int32)
        models.append((pts, v.copy(), feats, labels, int(m % 4)))
    return models


DATASET_CASES = {
    # every protocol, fixed number of points per model, rotation augmentation of points and of the normal features
    "sampled": dict(ctor=dict(numPoints=256, ptDropOut=1.0, pointFeatures=True, pointLabels=False, pointNormals=True,
                              useCategories=True, pointCategories=False, batchSize=4, allowedSamplings=[0, 1, 2, 3, 4],
                              augment=True, augmentMainAxis=1, augmentSmallRotations=True, augmentedFeatures=[0],
                              seed=7), epochs=2),
    # all points of every model (one acceptance pass), point drop-out, labels, per-point categories, a point budget
    "full": dict(ctor=dict(numPoints=0, ptDropOut=0.9, pointFeatures=False, pointLabels=True, pointNormals=True,
                           useCategories=True, pointCategories=True, batchSize=5, allowedSamplings=[0, 1, 2, 3],
                           maxPtsxBatch=2500, augment=True, augmentMainAxis=2, augmentSmallRotations=False, seed=11),
                 epochs=1),
    # more points than some models have (wrap-around passes), first-n selection
    "wrap": dict(ctor=dict(numPoints=1000, ptDropOut=0.8, pointFeatures=True, pointLabels=True, pointNormals=True,
                           useCategories=False, pointCategories=False, batchSize=3, allowedSamplings=[0, 1, 2, 3, 4],
                           uniformSelectFirst=True, seed=3), epochs=1),
}
