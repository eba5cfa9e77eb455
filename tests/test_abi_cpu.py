"""No-GPU checks of the drop-in boundary: the C-ABI shared library loads, exports every symbol include/mccnn_b200.h
declares (and nothing is bound in Python that the header does not declare), the host-side argument validation raises
the reference's error types, and the product refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mccnn_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mccnn_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libmccnn_b200.so does not export " + n
    # the ctypes table and the header describe the same surface
    assert sorted(_lib.SIGNATURES) == names


def test_header_cites_the_reference_interface():
    src = open(HEADER).read()
    for cite in ["aabb_gpu.cc:38-40", "sort_gpu.cc:119-131", "sort_gpu.cc:133-147", "find_neighbors.cc:45-56",
                 "find_neighbors.cc:63-78", "compute_pdf.cc:43-55", "poisson_sampling.cc:68-92",
                 "spatial_conv.cc:98-128"]:
        assert cite in src, cite


def test_constants_and_status_plumbing():
    from mccnn_b200._lib import lib, check
    assert lib.mc_get_block_size() == 8          # tf_ops/genCompileScript.py:20
    assert lib.mc_abi_version() >= 1
    assert lib.mc_build_grid_supported(32, 10) == 1 and lib.mc_build_grid_supported(32, 40) == 0
    assert lib.mc_sort_step1_workspace_ints(1000, 2, 10) >= 2000 + 2 * 1000
    # argument validation happens before any CUDA call: status 1 -> ValueError with a message
    st = lib.mc_find_neighbors_count(1, 0, None, 4, ctypes.c_float(-1.0), None, None, None, None, None, None, None,
                                     None, None, None)
    assert st == 1 and b"radius" in lib.mc_last_error()
    with pytest.raises(ValueError):
        check(st)
    st = lib.mc_spatial_conv(1, 1, 0, None, 12, 12, 0, None, 0, ctypes.c_float(0.1), 1, *([None] * 18))
    assert st == 1 and b"multiple of the block size" in lib.mc_last_error()


def test_op_api_surface_matches_the_reference_module():
    """tf_ops/MCConvModuleSrc:20-73 + genCompileScript.py:46-47: the 11 wrappers and get_block_size, same parameters."""
    import inspect
    import mccnn_b200.MCConvModule as M
    want = {
        "compute_aabb": ["inPts", "inBatchIds", "batchSize", "scaleInv"],
        "sort_points_step1": ["inPts", "inBatchIds", "aabbMin", "aabbMax", "batchSize", "cellSize", "scaleInv"],
        "sort_points_step2": ["inPts", "inBatchIds", "inFeatures", "keys", "indexs", "aabbMin", "aabbMax",
                              "batchSize", "cellSize", "scaleInv"],
        "sort_features": ["inFeatures", "indexs"],
        "sort_features_back": ["inFeatures", "indexs"],
        "transform_indexs": ["inIndexs", "inNewPositions"],
        "find_neighbors": ["inPts", "inBatchIds", "inPts2", "cellIndexs", "aabbMin", "aabbMax", "radius",
                           "batchSize", "scaleInv"],
        "compute_pdf": ["inPts", "inBatchIds", "aabbMin", "aabbMax", "startIndexs", "neighbors", "window", "radius",
                        "batchSize", "scaleInv"],
        "poisson_sampling": ["inPts", "inBatchIds", "cellIndexs", "aabbMin", "aabbMax", "radius", "batchSize",
                             "scaleInv"],
        "get_sampled_features": ["inSampledIndexs", "pInFeatures"],
        "spatial_conv": ["inPts", "inFeatures", "inBatchIds", "inPDFs", "inSamplePts", "neighStartIndexs",
                         "packedNeighs", "aabbMin", "aabbMax", "weights1", "weights2", "weightsOut", "biases1",
                         "biases2", "biasesOut", "numOutFeatures", "combin", "batchSize", "radius", "scaleInv", "avg"],
        "get_block_size": [],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(M, name)).parameters) == params, name
    assert M.get_block_size() == 8


def test_builder_surface_matches_the_reference_builder():
    import inspect
    from mccnn_b200.MCConvBuilder import PointHierarchy, ConvolutionBuilder
    assert list(inspect.signature(PointHierarchy.__init__).parameters)[1:] == [
        "inPoints", "inFeatures", "inBatchIds", "radiusList", "hierarchyName", "batchSize", "relativeRadius"]
    cc = list(inspect.signature(ConvolutionBuilder.create_convolution).parameters)[1:]
    assert cc == ["convName", "inPointHierarchy", "inPointLevel", "inFeatures", "inNumFeatures", "convRadius",
                  "outPointHierarchy", "outPointLevel", "multiFeatureConv", "outNumFeatures", "KDEWindow",
                  "relativeRadius", "usePDF", "useAVG"]
    b = ConvolutionBuilder.__init__
    assert list(inspect.signature(b).parameters)[1:7] == ["multiFeatureConvs", "KDEWindow", "relativeRadius", "usePDF",
                                                          "useAVG", "decayLossCollection"]


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing in the product imports the oracle."""
    import torch
    import mccnn_b200.MCConvModule as M
    pts = torch.zeros((8, 3))
    bids = torch.zeros((8, 1), dtype=torch.int32)
    with pytest.raises(RuntimeError):
        M.compute_aabb(pts, bids, 1, True)
    pkg = os.path.join(ROOT, "mccnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "libmccnn_ref" not in text, f
