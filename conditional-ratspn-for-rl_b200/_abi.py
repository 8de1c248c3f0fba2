"""ctypes binding of libspn_b200.so (the C ABI declared in include/spn_b200.h).

The library is the only compute backend of this package: there is no CPU path and no eager-PyTorch fallback.
``lib()`` raises if the shared object has not been built (run ``python __graft_entry__.py`` or
``python -m build_ext`` equivalent: ``build.py``), and every op wrapper raises if its tensors are not CUDA tensors.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPN_LIB_PATH: load another build of the same library (experiment builds of build.py; never a different implementation)
LIB_PATH = os.environ.get("SPN_LIB_PATH") or os.path.join(_HERE, "lib", "libspn_b200.so")

_c_int, _c_i64, _c_u64, _c_ptr = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p

# name -> argtypes, in the order of include/spn_b200.h
_P, _I, _L, _U, _F = _c_ptr, _c_int, _c_i64, _c_u64, ctypes.c_float
SIGNATURES = {
    "spn_version": [ctypes.c_char_p, _I],
    "spn_leaf_fwd": [_P, _L, _L, _L, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _L, _P],
    "spn_leaf_bwd_params": [_P, _L, _L, _L, _L, _P, _L, _L, _L, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P,
                            _P],
    "spn_leaf_bwd_x": [_P, _L, _L, _L, _L, _P, _L, _L, _L, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L,
                       _P],
    "spn_leaf_shared_fwd": [_P, _L, _L, _L, _L, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _L,
                            _P],
    "spn_leaf_shared_bwd_params": [_P, _L, _L, _L, _L, _P, _L, _L, _L, _L, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I,
                                   _I, _I, _P, _P, _P],
    "spn_leaf_workspace_floats": [_I, _I],
    "spn_leaf_transpose": [_P, _L, _L, _I, _I, _P, _P],
    "spn_leaf_tiled_supported": [_I, _I, _I],
    "spn_leaf_tiled_fwd": [_P, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _L, _P],
    "spn_leaf_tiled_bwd_params": [_P, _P, _L, _L, _L, _L, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "spn_leaf_tc_supported": [_I, _I, _I, _I],
    "spn_leaf_tc_image_bytes": [_I, _I, _I, _I],
    "spn_leaf_tc_fwd": [_P, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "spn_leaf_tc_bwd_params": [_P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "spn_sum_fwd": [_I, _P, _L, _L, _L, _L, _I, _I, _P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L,
                    _L, _L, _P],
    "spn_sum_bwd_w": [_I, _P, _L, _L, _L, _L, _I, _I, _P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P,
                      _L, _L, _L, _L, _P, _L, _P],
    "spn_sum_bwd_x": [_I, _P, _L, _L, _L, _L, _I, _I, _P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P,
                      _L, _L, _L, _L, _P, _L, _L, _L, _L, _P],
    "spn_sum_ps_supported": [_I, _I, _I],
    "spn_sum_ps_fwd": [_P, _L, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P],
    "spn_entropy_combine_fwd": [_P, _P, _P, _P, _L, _I, _I, _I, _I, _P, _P],
    "spn_entropy_combine_bwd": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "spn_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _P],
    "spn_sum_ps_fwd_raw": [_P, _L, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "spn_sum_ps_bwd_raw": [_P, _L, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _L, _P],
    "spn_sum_ps_bwd_supported": [_I, _I, _I],
    "spn_sum_ps_bwd": [_P, _L, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _L, _L, _P],
    "spn_xprod_fwd": [_P, _I, _I, _I, _I, _P, _P],
    "spn_xprod_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "spn_sample_sum": [_P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _I, _P, _P, _L, _L, _L, _L,
                       _I, _I, _I, _I, _I, _P, _I, _U, _U, _P, _P, _P],
    "spn_sum_tc_supported": [_I, _I, _I],
    "spn_sum_tc_padded_channels": [_I, _I],
    "spn_sum_tc_image_bytes": [_I, _I, _I, _I],
    "spn_sum_tc_prepare": [_P, _I, _I, _I, _I, _P, _P, _P],
    "spn_sum_tc_fhand_bytes": [_I, _I, _I, _I, _I],
    "spn_sum_tc_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "spn_sum_tc_hand_bytes": [_I, _I, _I, _I, _I],
    "spn_sum_tc_dw_workspace_floats": [_I, _I, _I, _I],
    "spn_sum_tc_bwd_x": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "spn_sum_tc_bwd_w": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "spn_debug_tc_aborts": [_P],
    "spn_debug_tc_prof": [_P, _I],
    "spn_root_shared_supported": [_I, _I, _I],
    "spn_root_shared_fwd": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P],
    "spn_root_shared_bwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "spn_sample_leaf": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _P, _P, _I, _U, _U, _P, _P, _P, _L, _L, _L,
                        _I, _I, _P, _P, _P],
    "spn_sample_sum_bwd": [_P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _P, _L, _L, _L, _I, _P, _P, _L, _L, _L,
                           _L, _I, _I, _I, _I, _I, _P, _U, _U, _P, _P, _I, _I, _P, _P, _P, _P],
    "spn_sample_leaf_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _U, _U, _P, _P, _P, _L, _L, _L, _I,
                            _I, _P, _I, _P, _P, _P, _P, _P],
}

_RETURNS_INT64 = {"spn_leaf_workspace_floats", "spn_leaf_tc_image_bytes", "spn_sum_tc_image_bytes", "spn_sum_tc_hand_bytes", "spn_sum_tc_fhand_bytes",
                  "spn_sum_tc_dw_workspace_floats"}

_lib = None
LAUNCHES = 0  # number of kernel-launching ABI calls made by this process (bench.py reports it)


class SpnAbiError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Fails loudly when it is missing (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpnAbiError(
                f"{LIB_PATH} is missing: build the CUDA library first (python __graft_entry__.py, or "
                f"python conditional-ratspn-for-rl_b200/build.py). This package has no CPU / eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int64 if name in _RETURNS_INT64 else ctypes.c_int
        _lib = handle
    return _lib


def version() -> str:
    buf = ctypes.create_string_buffer(128)
    lib().spn_version(buf, 128)
    return buf.value.decode()


def check(rc: int, what: str):
    if rc != 0:
        if rc <= -1000:
            raise SpnAbiError(f"{what}: CUDA error {-(rc + 1000)}")
        raise SpnAbiError(f"{what}: status {rc} (invalid argument / unsupported shape)")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    p = _StreamPtr(torch.cuda.current_stream(device).cuda_stream)
    if device is not None:
        device = torch.device(device)
        p.device_index = device.index if device.index is not None else torch.cuda.current_device()
    return p


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SpnAbiError(
                "ratspn_b200 runs on CUDA (sm_100a) only: got a tensor on "
                f"{t.device}. Move the model and its inputs to a B200 ('cuda'); there is no CPU path.")


class _StreamPtr(ctypes.c_void_p):
    """cudaStream_t handle that remembers which device it belongs to (see ``call``)."""
    device_index = None


def call(name, *args):
    """Invoke a kernel-launching entry point.  The ABI launches on the CURRENT device, so when the stream argument (last
    argument of every launching entry point, built by ``stream_ptr(tensor.device)``) belongs to another device --
    single-process multi-GPU use -- that device is made current around the call."""
    global LAUNCHES
    LAUNCHES += 1
    st = args[-1] if args else None
    idx = getattr(st, "device_index", None)
    if idx is not None and idx != torch.cuda.current_device():
        with torch.cuda.device(idx):
            check(getattr(lib(), name)(*args), name)
        return
    check(getattr(lib(), name)(*args), name)
