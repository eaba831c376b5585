"""dynamics_b200: B200-native (sm_100a) implementation of the Neural Physics Engine hot path of mbchang/dynamics.

The product is ``libnpe_cuda.so`` (hand-written CUDA behind a C ABI, ``include/npe_cuda.h``).  This package holds the
kernels (``csrc/``), the ctypes binding (``lib.py``), the host-side mirror of the reference's Lua ``model`` /
batcher / optimizer / rollout interfaces (``model.py``, ``sampler.py``, ``optim.py``, ``rollout.py``) and the Lua/FFI
shim a maintainer drops into the reference (``lua/``).  There is no CPU fallback: without the CUDA library and a GPU
every compute call raises.
"""
from .lib import NPEError, NPELibrary, load_library  # noqa: F401
from .model import NPEModel, ModelParams  # noqa: F401
