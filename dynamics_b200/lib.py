"""ctypes binding of libnpe_cuda.so (include/npe_cuda.h).  Mirrors what LuaJIT ``ffi.cdef`` binds in lua/npe_ffi.lua."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnpe_cuda.so")

NPE_NUM_PARAMS = 28222
OBJ_DIM = 22
INPUT_DIM = 44


class NPEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnpe_cuda error {code}: {msg}")
        self.code = code


class NPEConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("rnn_dim", C.c_int32), ("layers", C.c_int32),
                ("num_past", C.c_int32), ("num_future", C.c_int32), ("object_dim", C.c_int32), ("nbrhd", C.c_int32),
                ("nbrhdsize", C.c_float), ("max_ctx", C.c_int32), ("vlambda", C.c_float), ("lambda_", C.c_float),
                ("ball_zero_mode", C.c_int32), ("reserved", C.c_int32 * 4)]


_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_H = C.c_void_p

# name -> (restype, argtypes): every symbol include/npe_cuda.h declares
SIGNATURES = {
    "npe_abi_version": (C.c_int, []),
    "npe_default_config": (None, [C.POINTER(NPEConfig)]),
    "npe_create": (C.c_int, [C.POINTER(NPEConfig), C.POINTER(_H)]),
    "npe_destroy": (C.c_int, [_H]),
    "npe_last_error": (C.c_char_p, [_H]),
    "npe_param_count": (C.c_int, [_H]),
    "npe_params_set": (C.c_int, [_H, _fp, C.c_int32]),
    "npe_params_get": (C.c_int, [_H, _fp, C.c_int32]),
    "npe_grads_get": (C.c_int, [_H, _fp, C.c_int32]),
    "npe_grads_set": (C.c_int, [_H, _fp, C.c_int32]),
    "npe_batch_upload": (C.c_int, [_H, _fp, _fp, _fp, C.c_int32, C.c_int32]),
    "npe_scenes_upload": (C.c_int, [_H, _fp, _ip, C.c_int32, C.c_int32, C.c_int32]),
    "npe_scenes_upload_raw": (C.c_int, [_H, _fp, _ip, C.c_int32, C.c_int32, C.c_int32]),
    "npe_windows_upload": (C.c_int, [_H, _ip, _ip, C.c_int32]),
    "npe_select_windows": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "npe_foci_upload": (C.c_int, [_H, _ip, _ip, _ip, C.c_int32]),
    "npe_select_foci": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "npe_num_foci": (C.c_int, [_H]),
    "npe_ctx_slots": (C.c_int, [_H]),
    "npe_batch_download": (C.c_int, [_H, _fp, _fp]),
    "npe_build_pairs": (C.c_int, [_H, _ip, C.POINTER(C.c_uint8), _ip]),
    "npe_forward": (C.c_int, [_H, _fp, _fp]),
    "npe_forward_per_example": (C.c_int, [_H, _fp, _fp]),
    "npe_backward": (C.c_int, [_H, C.c_int32]),
    "npe_adam_step": (C.c_int, [_H, C.c_float, C.c_float, C.c_float, C.c_float]),
    "npe_rmsprop_step": (C.c_int, [_H, C.c_float, C.c_float, C.c_float]),
    "npe_optim_reset": (C.c_int, [_H, C.c_float]),
    "npe_train_step": (C.c_int, [_H, C.c_int32, C.c_float, C.c_int32, _fp]),
    "npe_train_step_async": (C.c_int, [_H, C.c_int32, C.c_float, C.c_int32, C.c_int32]),
    "npe_loss_wait": (C.c_int, [_H, C.c_int32, _fp]),
    "npe_train_batch_async": (C.c_int, [_H, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32]),
    "npe_train_steps": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _fp]),
    "npe_rollout": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, C.POINTER(C.c_double)]),
    "npe_rollout_slots": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, C.POINTER(C.c_double)]),
    "npe_rollout_resident": (C.c_int, [_H, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "npe_set_precision": (C.c_int, [_H, C.c_int32]),
    "npe_set_rollout_mode": (C.c_int, [_H, C.c_int32]),
    "npe_set_option": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "npe_comm_unique_id": (C.c_int, [C.c_void_p]),
    "npe_comm_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
    "npe_allreduce_grads": (C.c_int, [_H]),
    "npe_peer_export": (C.c_int, [_H, C.c_void_p]),
    "npe_peer_attach": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
    "npe_priority_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_float]),
    "npe_priority_update": (C.c_int, [_H, C.c_int32, _ip, _fp, C.c_int32]),
    "npe_priority_sample": (C.c_int, [_H, C.c_int32, C.c_float, C.c_uint64, C.c_int32, _ip]),
    "npe_priority_hardest": (C.c_int, [_H, C.c_int32, _fp, _ip]),
    "npe_sync": (C.c_int, [_H]),
    "npe_host_alloc": (C.c_void_p, [C.c_size_t]),
    "npe_host_free": (None, [C.c_void_p]),
    "npe_step_trace": (C.c_int, [_H, C.POINTER(C.c_uint64), C.c_int32]),
    "npe_profile_select": (C.c_int, [_H, C.c_int32]),
    "npe_profile_period": (C.c_int, [_H, C.c_int32]),
    "npe_profile_read": (C.c_int, [_H, _ip, C.POINTER(C.c_double)]),
    "npe_event_record": (C.c_int, [_H, C.c_int32]),
    "npe_event_elapsed_ms": (C.c_int, [_H, C.c_int32, C.c_int32, _fp]),
    "npe_launch_count": (C.c_int64, [_H]),
    "npe_debug_read": (C.c_int, [_H, C.c_int32, _fp, C.c_int64]),
    "npe_nbr_threshold_sq": (C.c_float, [C.c_float]),
    "npe_batch_stats": (C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


class NPELibrary:
    """Loaded libnpe_cuda.so with typed entry points.  Raises if the library is missing: there is no fallback."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -m dynamics_b200.build` (nvcc, sm_100a). "
                "dynamics_b200 has no CPU or PyTorch fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def default_config(self, **over):
        cfg = NPEConfig()
        self.npe_default_config(C.byref(cfg))
        for k, v in over.items():
            setattr(cfg, "lambda_" if k == "lambda" else k, v)
        return cfg


_LIB = None


def load_library(path=LIB_PATH):
    global _LIB
    if _LIB is None or _LIB.path != path:
        _LIB = NPELibrary(path)
    return _LIB


OPTIONS = {"train_fwd_tc": 0, "train_stash": 1, "bwd_tc": 2, "peer_timeout_ms": 3, "small_batch_max": 4, "dec_stacked": 5}
KERNEL_IDS = {"build_pairs": 0, "forward": 1, "dec_backward": 2, "pair_backward": 3, "reduce": 4, "optim": 5,
              "rollout": 6, "targets": 7, "step_fused": 8, "wgrad": 9}


def pinned_array(lib, shape, dtype=np.float32):
    """numpy view over page-locked host memory owned by the library (npe_host_alloc)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.npe_host_alloc(max(n, 16))
    if not ptr:
        raise MemoryError("npe_host_alloc failed")
    buf = (C.c_char * max(n, 16)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, ptr


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fptr(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None
