"""Network topologies of the reference (models/*.py) expressed against mccnn_b200's builder API."""
from . import MCClass, MCClassH, MCClassS, MCNorm, MCNormS, MCSeg, MCSegScanNet  # noqa: F401
