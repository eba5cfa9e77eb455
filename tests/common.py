"""Helpers shared by the parity tests: the seeded synthetic inputs (re-exported from the repository's synthetic.py),
error measures, and the small model / data-set cases."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import make_cloud, make_batch, make_conv_weights, named_variable  # noqa: E402,F401


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
