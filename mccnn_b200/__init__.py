"""mccnn_b200 -- B200-native (sm_100a) Monte-Carlo convolution forward path of viscom-ulm/MCCNN.

  MCConvModule     the reference's Python op API (tf_ops/MCConvModuleSrc) on torch CUDA tensors
  MCConvBuilder    PointHierarchy / ConvolutionBuilder (utils/MCConvBuilder.py)
  MCNetworkUtils   dense glue (utils/MCNetworkUtils.py), torch ops
  models           network topologies (models/MCClassS.py ...)
  runtime          exact / record / planned execution modes (sync-free, CUDA-graph capturable)
  engine           recorded + graph-captured forward, data-parallel sharding over GPUs
  ops, _lib        raw ctypes bindings of libmccnn_b200.so (include/mccnn_b200.h)

Importing the package loads the shared library and fails loudly if it is missing: there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is not built)
from . import ops, runtime, variables, MCConvModule, MCConvBuilder, MCNetworkUtils  # noqa: F401
from . import models  # noqa: F401

__version__ = "0.1.0"
