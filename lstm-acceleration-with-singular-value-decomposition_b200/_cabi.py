"""ctypes binding of libsvdlstm.so (the C-ABI declared in include/svdlstm.h).

There is NO CPU fallback anywhere in this package: if the shared library is missing, or no CUDA
device is visible, every compute entry point raises.  PyTorch is used only for device memory,
streams and torch.distributed.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVDLSTM_LIB") or os.path.join(_HERE, "libsvdlstm.so")   # SVDLSTM_LIB: e.g. the timeline debug build

RETURN_SEQUENCES = 1
GO_BACKWARDS = 2
TIME_MAJOR = 4
ZERO_OUTPUT_FOR_MASK = 8

ENGINE_AUTO = 0
ENGINE_GENERAL = 1
ENGINE_WAVEFRONT = 2
ENGINE_TC = 3          # tcgen05 tensor-core engine: FP16 operands, FP32 accumulation / cell state (reduced precision)
ENGINE_FP32 = 4        # strict FP32: wavefront when the model fits it, else general (what "auto" was before the regime switch)
TC_MIN_BATCH = 128     # "auto" picks the tensor-core engine from this batch size on (include/svdlstm.h)
TC_MIN_UNITS = 64
ENGINE_NAMES = {"auto": 0, "general": 1, "wavefront": 2, "tc": 3, "tc_f16": 3, "fp32": 4, None: 0}

EXPORTS = [
    "svdlstm_create", "svdlstm_destroy", "svdlstm_set_full_weights", "svdlstm_set_singular_weights",
    "svdlstm_set_reduced_weights", "svdlstm_set_dense_top", "svdlstm_forward", "svdlstm_forward_streamed_input", "svdlstm_last_launches",
    "svdlstm_last_engine", "svdlstm_count_weights", "svdlstm_host_alloc", "svdlstm_host_free", "svdlstm_svd_jacobi_batched",
    "svdlstm_reduce_factors", "svdlstm_reduce_factors_batched", "svdlstm_scaled_matmul", "svdlstm_penalties", "svdlstm_sweep_sse", "svdlstm_last_error",
    "svdlstm_version", "svdlstm_stream_open", "svdlstm_stream_step", "svdlstm_stream_run", "svdlstm_stream_reset",
    "svdlstm_stream_state", "svdlstm_stream_launches", "svdlstm_stream_close",
    "svdlstm_trainer_create", "svdlstm_trainer_destroy", "svdlstm_trainer_num_params", "svdlstm_trainer_layout",
    "svdlstm_trainer_gradients", "svdlstm_trainer_regularizers", "svdlstm_trainer_adam",
]


class PenaltyItem(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("ld", ctypes.c_int32), ("gram", ctypes.c_int32), ("columns", ctypes.c_int32)]


class ReduceItem(ctypes.Structure):
    _fields_ = [("U_r", ctypes.c_void_p), ("S_r", ctypes.c_void_p), ("V_r", ctypes.c_void_p), ("B", ctypes.c_void_p),
                ("C", ctypes.c_void_p), ("pivot_ratio", ctypes.c_void_p), ("ldu", ctypes.c_int32), ("ldv", ctypes.c_int32),
                ("m", ctypes.c_int32), ("r", ctypes.c_int32), ("n", ctypes.c_int32)]


_lib = None
# kernels launched through this binding since import (bench.py's `gpu_launches` reads it)
launch_counter = 0


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsvdlstm.so is not built (%s missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python lstm-acceleration-with-singular-value-decomposition_b200/build.py`. "
            "This package has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cip = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)
    L.svdlstm_create.argtypes = [ctypes.POINTER(vp), ci, ci, cip]
    L.svdlstm_create.restype = ci
    L.svdlstm_destroy.argtypes = [vp]
    L.svdlstm_destroy.restype = None
    L.svdlstm_set_full_weights.argtypes = [vp, ci, vp, vp, vp]
    L.svdlstm_set_full_weights.restype = ci
    L.svdlstm_set_singular_weights.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ci, ci]
    L.svdlstm_set_singular_weights.restype = ci
    L.svdlstm_set_reduced_weights.argtypes = [vp, ci, ci, ctypes.POINTER(vp), cip]
    L.svdlstm_set_reduced_weights.restype = ci
    L.svdlstm_set_dense_top.argtypes = [vp, vp, vp, ci]
    L.svdlstm_set_dense_top.restype = ci
    L.svdlstm_forward.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]
    L.svdlstm_forward.restype = ci
    L.svdlstm_forward_streamed_input.argtypes = [vp, vp, vp, ci, ci, vp, ci, vp, vp]
    L.svdlstm_forward_streamed_input.restype = ci
    L.svdlstm_last_launches.argtypes = [vp]
    L.svdlstm_last_launches.restype = ci
    L.svdlstm_last_engine.argtypes = [vp]
    L.svdlstm_last_engine.restype = ci
    L.svdlstm_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t, ci]
    L.svdlstm_host_alloc.restype = ci
    L.svdlstm_host_free.argtypes = [vp]
    L.svdlstm_host_free.restype = ci
    L.svdlstm_count_weights.argtypes = [vp]
    L.svdlstm_count_weights.restype = ctypes.c_int64
    L.svdlstm_svd_jacobi_batched.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp]
    L.svdlstm_svd_jacobi_batched.restype = ci
    L.svdlstm_reduce_factors.argtypes = [vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp]
    L.svdlstm_reduce_factors.restype = ci
    L.svdlstm_reduce_factors_batched.argtypes = [ctypes.POINTER(ReduceItem), ci, vp]
    L.svdlstm_reduce_factors_batched.restype = ci
    L.svdlstm_scaled_matmul.argtypes = [vp, ci, vp, vp, ci, vp, ci, ci, ci, vp, ci, vp]
    L.svdlstm_scaled_matmul.restype = ci
    L.svdlstm_penalties.argtypes = [ctypes.POINTER(PenaltyItem), ci, vp, vp]
    L.svdlstm_penalties.restype = ci
    L.svdlstm_sweep_sse.argtypes = [vp, vp, ci, ctypes.c_int64, vp, vp]
    L.svdlstm_sweep_sse.restype = ci
    L.svdlstm_stream_open.argtypes = [vp, ci, ctypes.POINTER(vp)]
    L.svdlstm_stream_open.restype = ci
    L.svdlstm_stream_step.argtypes = [vp, vp, vp]
    L.svdlstm_stream_step.restype = ci
    L.svdlstm_stream_run.argtypes = [vp, vp, ci, ctypes.c_double, vp, vp]
    L.svdlstm_stream_run.restype = ci
    L.svdlstm_stream_reset.argtypes = [vp, vp, vp]
    L.svdlstm_stream_reset.restype = ci
    L.svdlstm_stream_state.argtypes = [vp, vp, vp]
    L.svdlstm_stream_state.restype = ci
    L.svdlstm_stream_launches.argtypes = [vp]
    L.svdlstm_stream_launches.restype = ci
    L.svdlstm_stream_close.argtypes = [vp]
    L.svdlstm_stream_close.restype = ci
    i64p = ctypes.POINTER(ctypes.c_int64)
    L.svdlstm_trainer_create.argtypes = [vp, cip, ctypes.POINTER(vp)]
    L.svdlstm_trainer_create.restype = ci
    L.svdlstm_trainer_destroy.argtypes = [vp]
    L.svdlstm_trainer_destroy.restype = None
    L.svdlstm_trainer_num_params.argtypes = [vp]
    L.svdlstm_trainer_num_params.restype = ctypes.c_int64
    L.svdlstm_trainer_layout.argtypes = [vp, i64p]
    L.svdlstm_trainer_layout.restype = ci
    L.svdlstm_trainer_gradients.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
    L.svdlstm_trainer_gradients.restype = ci
    L.svdlstm_trainer_regularizers.argtypes = [vp, cip, cip, ctypes.POINTER(ctypes.c_float), ci, vp, vp, vp]
    L.svdlstm_trainer_regularizers.restype = ci
    L.svdlstm_trainer_adam.argtypes = [vp, vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, vp]
    L.svdlstm_trainer_adam.restype = ci
    L.svdlstm_last_error.argtypes = []
    L.svdlstm_last_error.restype = ctypes.c_char_p
    L.svdlstm_version.argtypes = []
    L.svdlstm_version.restype = ctypes.c_char_p
    _lib = L
    return L


def check(rc: int) -> None:
    """0 ok; <0 argument/shape error -> ValueError (as Keras raises on bad set_weights);
    >0 cudaError_t -> RuntimeError."""
    if rc == 0:
        return
    msg = lib().svdlstm_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("svdlstm: no CUDA device visible; this package has no CPU fallback "
                           "(the numpy oracle under oracle/ is test infrastructure only)")
    return torch.device("cuda", torch.cuda.current_device())


def dev_tensor(a, device: Optional[torch.device] = None) -> torch.Tensor:
    """numpy / torch -> contiguous float32 CUDA tensor."""
    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float32)), device=device)


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def int_array(vals: Sequence[int]):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def scaled_matmul(A: torch.Tensor, B: torch.Tensor, scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                  k: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = (A[:, :k] * scale[:k]) @ B[:k] + bias on device (K5; float32, row strides honoured, last dim contiguous)."""
    if A.stride(-1) != 1:
        A = A.contiguous()
    if B.stride(-1) != 1:
        B = B.contiguous()
    m, n = int(A.shape[0]), int(B.shape[1])
    k = int(A.shape[1]) if k is None else int(k)
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=A.device)
    if out.stride(-1) != 1:
        raise ValueError("scaled_matmul: out must have a contiguous last dimension")
    check(lib().svdlstm_scaled_matmul(ptr(A), A.stride(0), ptr(scale), ptr(B), B.stride(0), ptr(bias), m, k, n, ptr(out), out.stride(0),
                                      cur_stream()))
    add_launches(1)
    return out


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; freed when the last tensor viewing it dies."""

    def __init__(self, nbytes: int, write_combined: bool):
        self.p = ctypes.c_void_p()
        check(lib().svdlstm_host_alloc(ctypes.byref(self.p), nbytes, 1 if write_combined else 0))
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.p.value:
                lib().svdlstm_host_free(self.p)
                self.p = ctypes.c_void_p()
        except Exception:
            pass


_pinned_pool = {}      # nbytes -> free _PinnedBlocks


def pooled_pinned_array(shape) -> np.ndarray:
    """float32 numpy array in page-locked memory taken from a pool; the block returns to the pool when the array (and every view
    of it) has died.  ``predict`` hands its result out this way: the device->host copy lands in memory the caller then OWNS -- no
    second copy out of a staging buffer -- and page-locking (milliseconds per allocation) is paid once per size."""
    import weakref
    n = int(np.prod(shape))
    nbytes = max(4 * n, 4)
    if nbytes not in _pinned_pool:
        # first result of this size: page-lock TWO blocks now (tens of ms each for a 16 MB result) -- `y = model.predict(x)` in a
        # loop holds the previous result while the next one is produced, so a second block is needed by the second call anyway
        _pinned_pool[nbytes] = [_PinnedBlock(nbytes, False)]
    free = _pinned_pool[nbytes]
    blk = free.pop() if free else _PinnedBlock(nbytes, False)
    buf = (ctypes.c_float * n).from_address(blk.p.value)
    weakref.finalize(buf, free.append, blk)      # (the callback keeps the block alive; it runs when the array's base is collected)
    return np.ctypeslib.as_array(buf).reshape(shape)


def pinned_empty(shape, write_combined: bool = False) -> torch.Tensor:
    """float32 CPU tensor in page-locked host memory (a staging buffer for ``predict`` / ``predict_async``).
    ``write_combined=True`` -> cudaHostAllocWriteCombined: fill it once from the CPU, let the GPU read it."""
    n = int(np.prod(shape))
    blk = _PinnedBlock(max(4 * n, 4), write_combined)
    buf = (ctypes.c_float * n).from_address(blk.p.value)
    buf._svdlstm_block = blk        # lifetime: torch storage -> ndarray -> ctypes view -> block (views of the tensor keep it alive too)
    return torch.from_numpy(np.ctypeslib.as_array(buf).reshape(shape))


def add_launches(n: int) -> None:
    global launch_counter
    launch_counter += int(n)
