"""ctypes binding of libmccnn_b200.so (include/mccnn_b200.h).

The product has exactly one compute path: the sm_100a kernels in mccnn_b200/csrc.  If the shared library is
missing or fails to load this module raises -- there is no CPU or eager-PyTorch fallback.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libmccnn_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_ll = ctypes.c_longlong
_ptr = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mccnn_b200.h one to one
SIGNATURES = {
    "mc_last_error": (ctypes.c_char_p, []),
    "mc_abi_version": (_c_int, []),
    "mc_get_block_size": (_c_int, []),
    "mc_device_info": (_c_int, [_ptr, _ptr, _ptr]),
    "mc_launch_count": (_c_ll, []),
    "mc_compute_aabb": (_c_int, [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_determine_num_cells": (_c_int, [_c_int, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_verify_num_cells": (_c_int, [_c_float, _ptr, _ptr, _c_int, _ptr, _ptr]),
    "mc_sort_step1_workspace_ints": (_c_ll, [_c_int, _c_int, _c_int]),
    "mc_sort_points_step1": (_c_int, [_c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_sort_points_step2": (_c_int, [_c_int, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                      _ptr, _ptr, _ptr, _ptr]),
    "mc_build_grid_supported": (_c_int, [_c_int, _c_int]),
    "mc_build_grid_workspace_ints": (_c_ll, [_c_int]),
    "mc_build_grid": (_c_int, [_c_int, _ptr, _c_int, _c_int, _c_int] + [_ptr] * 14),
    "mc_sort_features": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr]),
    "mc_sort_features_back": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr]),
    "mc_sort_features_affine": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr]),
    "mc_transform_indexs": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_find_neighbors_workspace_ints": (_c_ll, [_c_int]),
    "mc_find_neighbors_count": (_c_int, [_c_int, _c_int, _ptr, _c_int, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                         _ptr, _ptr, _ptr, _ptr]),
    "mc_find_neighbors_pack": (_c_int, [_c_int, _c_int, _ptr, _c_int, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                        _ptr, _ptr, _ptr, _c_int, _ptr, _ptr]),
    "mc_find_neighbors_fused": (_c_int, [_c_int, _c_int, _ptr, _c_int, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                         _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr]),
    "mc_find_neighbors_pdf_fused": (_c_int, [_c_int, _c_float, _c_int, _ptr, _c_int, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr,
                                             _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr]),
    "mc_compute_pdf": (_c_int, [_c_int, _c_float, _c_int, _ptr, _c_int, _ptr, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr,
                                _ptr, _ptr, _ptr]),
    "mc_poisson_workspace_ints": (_c_ll, [_c_int, _c_int]),
    "mc_poisson_sampling": (_c_int, [_c_int, _c_float, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr,
                                     _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_get_sampled_features": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr]),
    "mc_spatial_conv_workspace_floats": (_c_ll, [_c_int, _c_int, _c_int, _c_int]),
    "mc_spatial_conv": (_c_int, [_c_int, _c_int, _c_int, _ptr, _c_int, _c_int, _c_int, _ptr, _c_int, _c_float, _c_int] +
                        [_ptr] * 18),
    "mc_scale_shift_act": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr]),
    "mc_dense_layer": (_c_int, [_c_int, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr,
                                _c_int, _ptr, _ptr]),
    "mc_mlp_rows": (_c_int, [_c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                             _ptr]),
    "mc_split_tf32": (_c_int, [_c_ll, _ptr, _ptr, _ptr, _ptr]),
    "mc_dense_fewrows_workspace_floats": (_c_ll, [_c_int, _c_int, _c_int]),
    "mc_dense_layer_fewrows": (_c_int, [_c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int,
                                        _ptr, _ptr, _ptr]),
    "mc_dense_layer_tc_supported": (_c_int, [_c_int, _c_int]),
    "mc_dense_layer_tc": (_c_int, [_c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr,
                                   _c_int, _ptr, _ptr]),
    "mc_spatial_conv_grad_workspace_floats": (_c_ll, [_c_int, _c_int, _c_int, _c_int]),
    "mc_spatial_conv_grad": (_c_int, [_c_int] * 8 + [_c_float, _c_int] + [_ptr] * 25),
    "mc_sort_points_step2_grad": (_c_int, [_c_int, _c_int] + [_ptr] * 6),
    "mc_get_sampled_features_grad": (_c_int, [_c_int, _c_int, _c_int] + [_ptr] * 4),
    "mc_allgather_rows_p2p_buffer_floats": (_c_ll, [_c_int, _c_int]),
    "mc_allgather_rows_p2p": (_c_int, [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_peer_exchange_buffer_floats": (_c_ll, [_c_int, _c_int, _c_int, _c_int]),
    "mc_peer_put": (_c_int, [_c_int, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "mc_peer_collect": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr]),
}


def build(verbose=False):
    """Compile mccnn_b200/csrc for sm_100a into mccnn_b200/_C/libmccnn_b200.so (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libmccnn_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "mccnn_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C mccnn_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(status, what=""):
    """Turn a C-ABI status into the exception the reference's glue would raise (InvalidArgument -> ValueError)."""
    if status == 0:
        return
    msg = lib.mc_last_error().decode("utf-8", "replace")
    if status == 1:
        raise ValueError(msg or what)
    raise RuntimeError(msg or what)
