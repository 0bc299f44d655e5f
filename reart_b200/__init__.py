"""reart_b200 -- B200-native (sm_100a) implementation of reart's per-iteration energy evaluation.

Host side mirrors the reference's Python interfaces for this path (utils/chamfer.py,
networks/loss.py, networks/model.py, screw_se3, utils/kinematic_utils.fk); the arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``libreart_b200.so`` (include/reart_b200.h).
"""
from ._lib import ReartError, SO_PATH, lib  # noqa: F401

__version__ = "0.1.0"
