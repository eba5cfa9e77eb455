"""Network topologies of the reference (models/*.py) expressed against mccnn_b200's builder API."""
from . import MCClassS  # noqa: F401
