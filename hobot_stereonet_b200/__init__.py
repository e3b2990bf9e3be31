"""B200-native StereoNet inference path (drop-in for the DnnNode::Run span of hobot_stereonet).

Only the hot path lives here: `csrc/` (sm_100a kernels + C ABI, built into lib/libsnb200.so),
`capi.py` (ctypes binding of include/snb200.h) and `host/` (host-side mirror of the reference
node's interface).  Nothing in this package imports `oracle/`.
"""
from . import capi  # noqa: F401
from .capi import Model, Pool, SnbError  # noqa: F401
