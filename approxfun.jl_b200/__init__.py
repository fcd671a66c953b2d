"""approxfun.jl_b200 -- B200-native (sm_100a) drop-in for ApproxFun.jl's transform / evaluate / sample hot path.

The directory name follows the reference repo ("ApproxFun.jl"), which is not a valid Python identifier:
import it through the repo-root shim `import approxfun_b200` (see approxfun_b200.py).
All arithmetic runs in libapproxfun_b200.so (hand-written CUDA behind a C ABI, include/approxfun_b200.h);
there is no CPU fallback.
"""
from . import _lib
from ._lib import AfbError, DimensionMismatch, LIB_PATH
from .api import *  # noqa: F401,F403
from . import api
from . import device
from .device import DeviceArray, DeviceFun
