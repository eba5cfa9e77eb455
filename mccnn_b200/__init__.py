"""mccnn_b200 -- B200-native (sm_100a) Monte-Carlo convolution forward path of viscom-ulm/MCCNN.

  MCConvModule     the reference's Python op API (tf_ops/MCConvModuleSrc) on torch CUDA tensors
  MCConvBuilder    PointHierarchy / ConvolutionBuilder (utils/MCConvBuilder.py)
  MCNetworkUtils   dense glue (utils/MCNetworkUtils.py): own fused kernels at inference, torch ops in training mode
  DataSet          batch assembly and sampling protocols on the device (utils/DataSet.py)
  models           network topologies (models/MCClassS.py ...)
  runtime          exact / record / planned execution modes (sync-free, CUDA-graph capturable)
  engine           recorded + graph-captured forward, data-parallel sharding over GPUs
  ops, _lib        raw ctypes bindings of libmccnn_b200.so (include/mccnn_b200.h)

Importing the package loads the shared library and fails loudly if it is missing: there is no CPU fallback.
"""
import os as _os

# engine.PipelinedForward keeps several forwards in flight on (depth x 3) streams; with the driver's default of 8
# hardware work queues, streams alias onto the same queue and serialise.  Only effective if set before the CUDA
# context is created; an explicit setting by the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402  (raises ImportError when the CUDA library is not built)
from . import ops, runtime, variables, MCConvModule, MCConvBuilder, MCNetworkUtils  # noqa: F401,E402
from . import models  # noqa: F401,E402

__version__ = "0.1.0"
