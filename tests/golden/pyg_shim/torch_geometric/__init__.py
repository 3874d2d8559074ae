"""Minimal stand-in for the three torch_geometric symbols the reference imports
(`tsp/net.py:5,15,21`, `tsp/utils.py:2`).  torch_geometric is not installable in this
image; this shim exists ONLY so `tests/golden/make_golden.py` can import the unmodified
reference files from /root/reference in the build container.  Test infrastructure."""
from . import nn, data  # noqa: F401
