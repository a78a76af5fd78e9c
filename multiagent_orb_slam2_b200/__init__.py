"""B200-native ORB frontend (extraction + binary descriptor matching) behind the ORBextractor /
ORBmatcher interface of andresenwc/MultiAgent_ORB_SLAM2. Hand-written sm_100a CUDA kernels behind
a C ABI (include/orb_b200.h); no CPU fallback."""
from ._lib import OrbError, lib  # noqa: F401
from .extractor import KP_DTYPE, ORBextractor  # noqa: F401
from .matcher import ORBmatcher  # noqa: F401
from . import guided  # noqa: F401,E402  (attaches the guided searches a-11 ... a-15 to ORBmatcher)
