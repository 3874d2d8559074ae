"""deepaco_b200 -- B200-native (sm_100a) rollout engine behind the DeepACO `ACO` / `Net` class surface.

Sub-packages mirror the reference's problem directories: `deepaco_b200.tsp`, `deepaco_b200.tsp_nls`,
`deepaco_b200.cvrp`.  The compute lives in libdeepaco_b200.so (C ABI: include/deepaco_b200.h).
"""
from ._lib import DeepAcoError, LIB_PATH, lib  # noqa: F401

__version__ = "0.1.0"
