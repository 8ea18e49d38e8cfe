"""ctypes binding of include/caz_b200.h.

The native library is the product: if it is missing, or no sm_100 device is visible, every entry
point raises -- there is no CPU fallback and nothing here imports ``oracle``.
"""
import ctypes
import os
import weakref

import numpy as np

from . import build as _build

CAZ_OK = 0
PRECISION_BF16 = 0
PRECISION_FP32 = 1
PRECISION_FP16 = 2
PRECISION_BF16_PURE = 3
EVAL_MASKED_SOFTMAX = 1
RECORD_BYTES = 68
STAGES = 4
STAGE_NAMES = ("pack", "stem", "tower", "heads")

SYMBOLS = (
    "caz_create", "caz_reload", "caz_destroy", "caz_eval", "caz_eval_device", "caz_eval_records", "caz_sync",
    "caz_encode", "caz_push_priors", "caz_puct_select", "caz_puct_select_device", "caz_last_error", "caz_abi_version", "caz_max_batch",
    "caz_precision", "caz_stream", "caz_kernel_launches", "caz_weight_bytes", "caz_inexact_weights", "caz_profile", "caz_profile_read",
    "caz_host_alloc", "caz_host_free", "caz_debug_conv_layer", "caz_set_small_batch_limit", "caz_set_fused_tower", "caz_debug_tower_trace", "caz_debug_sb_trace", "caz_submit", "caz_submit_records", "caz_submit_records_priors", "caz_collect",
    "caz_train_create", "caz_train_destroy", "caz_train_last_error", "caz_train_param_count", "caz_train_num_params", "caz_train_steps",
    "caz_train_forward_backward", "caz_train_forward_backward_begin", "caz_train_forward_backward_end", "caz_train_wait_gradients",
    "caz_train_gradient_split", "caz_train_apply", "caz_train_read", "caz_train_read_outputs", "caz_train_grad_device_ptr",
)


class CazError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"caz_b200 error {code}: {message}")
        self.code = code


class Arch(ctypes.Structure):
    _fields_ = [("input_planes", ctypes.c_int32), ("stem_kernel", ctypes.c_int32), ("filters", ctypes.c_int32),
                ("res_blocks", ctypes.c_int32), ("block_kernel", ctypes.c_int32), ("policy_filters", ctypes.c_int32),
                ("value_filters", ctypes.c_int32), ("value_fc", ctypes.c_int32), ("n_labels", ctypes.c_int32),
                ("bn_epsilon", ctypes.c_float)]


class Tensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("numel", ctypes.c_int64)]


_lib = None


def library_path():
    return _build.LIB_PATH


def load():
    """dlopen the in-tree library (building it first if nvcc is present and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        try:
            _build.build_native()
        except Exception as exc:  # noqa: BLE001 - re-raised with context
            raise ImportError(f"native library {path} is missing and could not be built: {exc}") from exc
    lib = ctypes.CDLL(path)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{path} does not export {missing}")
    vp, i32, i64, u32, f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_double
    lib.caz_create.argtypes = [ctypes.POINTER(Arch), ctypes.POINTER(Tensor), i32, i32, i32, i32, ctypes.POINTER(vp)]
    lib.caz_reload.argtypes = [vp, ctypes.POINTER(Tensor), i32]
    lib.caz_destroy.argtypes = [vp]
    lib.caz_destroy.restype = None
    lib.caz_eval.argtypes = [vp, vp, i32, vp, vp, vp, u32]
    lib.caz_eval_device.argtypes = [vp, vp, i32, vp, vp, vp, u32]
    lib.caz_eval_records.argtypes = [vp, vp, i32, vp, vp, vp, u32]
    lib.caz_sync.argtypes = [vp]
    lib.caz_puct_select_device.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, f64, f64, vp]
    lib.caz_submit.argtypes = [vp, vp, i32, vp, vp, vp, u32, ctypes.POINTER(i32)]
    lib.caz_submit_records.argtypes = [vp, vp, i32, vp, vp, vp, u32, ctypes.POINTER(i32)]
    lib.caz_submit_records_priors.argtypes = [vp, vp, i32, vp, vp, vp, vp, ctypes.POINTER(i32)]
    lib.caz_collect.argtypes = [vp, i32]
    lib.caz_encode.argtypes = [vp, vp, i32, vp]
    lib.caz_push_priors.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    lib.caz_puct_select.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, f64, f64, vp]
    lib.caz_last_error.argtypes = [vp]
    lib.caz_last_error.restype = ctypes.c_char_p
    lib.caz_abi_version.restype = i32
    lib.caz_max_batch.argtypes = [vp]
    lib.caz_precision.argtypes = [vp]
    lib.caz_stream.argtypes = [vp]
    lib.caz_stream.restype = vp
    lib.caz_kernel_launches.argtypes = [vp]
    lib.caz_kernel_launches.restype = i64
    lib.caz_weight_bytes.argtypes = [vp]
    lib.caz_weight_bytes.restype = i64
    lib.caz_inexact_weights.argtypes = [vp]
    lib.caz_inexact_weights.restype = i64
    lib.caz_profile.argtypes = [vp, i32]
    lib.caz_profile_read.argtypes = [vp, vp, vp, i32]
    lib.caz_set_small_batch_limit.argtypes = [vp, i32]
    lib.caz_set_fused_tower.argtypes = [vp, i32]
    lib.caz_debug_sb_trace.argtypes = [vp, i32, vp, i32]
    lib.caz_debug_tower_trace.argtypes = [vp, i32, vp, i32]
    lib.caz_debug_conv_layer.argtypes = [vp, i32, vp, vp, i32, vp]
    lib.caz_host_alloc.argtypes = [i64]
    lib.caz_host_alloc.restype = vp
    lib.caz_host_free.argtypes = [vp]
    lib.caz_host_free.restype = None
    _lib = lib
    return lib


def check(handle, rc):
    if rc != CAZ_OK:
        msg = load().caz_last_error(handle)
        raise CazError(rc, msg.decode("utf-8", "replace") if msg else "unknown")


_PTR_CACHE = {}


def ptr(a):
    """c_void_p of an array's data.  ndarray.ctypes builds a helper object on every access (~1 us, four of them per call on
    the latency path): the pointer of an array that does not own its memory -- a PinnedArray view, which cannot be resized in
    place -- is remembered by object identity (checked through a weak reference)."""
    if a is None:
        return None
    hit = _PTR_CACHE.get(id(a))
    if hit is not None and hit[0]() is a:
        return hit[1]
    p = ctypes.c_void_p(a.ctypes.data)
    if not a.flags.owndata:
        if len(_PTR_CACHE) >= 64:
            _PTR_CACHE.clear()
        _PTR_CACHE[id(a)] = (weakref.ref(a), p)
    return p


def as_array(a, dtype, shape=None):
    """C-contiguous view/copy of ``a`` with the given dtype (what np.asarray(..., dtype) gives the reference)."""
    out = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(out.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(out.shape)}")
    return out


class PinnedArray:
    """A numpy array over cudaHostAlloc memory: ``.array`` is valid until ``close()`` / garbage collection
    of this owner object.  Keep the owner alive for as long as the array is in use."""

    def __init__(self, shape, dtype):
        lib = load()
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        self._nbytes = max(count * dtype.itemsize, 1)
        self._ptr = lib.caz_host_alloc(self._nbytes)
        if not self._ptr:
            raise CazError(2, "cudaHostAlloc failed")
        buf = (ctypes.c_uint8 * self._nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def close(self):
        if self._ptr and _lib is not None:
            self.array = None
            _lib.caz_host_free(self._ptr)
        self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
