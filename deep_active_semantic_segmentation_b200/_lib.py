"""ctypes binding of libdas_b200.so - the only way the Python layer reaches the CUDA kernels.

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised (DasError / OSError).  Prototypes mirror include/das_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# DAS_B200_LIB: load another build of the library (A/B measurements of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("DAS_B200_LIB") or os.path.join(PKG, "libdas_b200.so")

ABI_VERSION = 4
MC_VOTES = 1
MC_PROBS = 2
MC_SINGLE_SHOT = 4
N_SCORES = 6
SCORE_INDEX = {"vote_entropy": 0, "pred_entropy": 1, "bald": 2, "confidence": 3, "margin": 4, "expected_entropy": 5}
ACC_INDEX = {"wrong_count": 0, "p0_sum": 1, "not_argmax_sum": 2, "unsure_mean": 3, "valid_count": 4}
MAX_CLASSES = 32
MAX_PASSES = 255
MAX_PASS_GROUP = 32
TOPK_MAX_K = 4096
#: das_handle options (include/das_b200.h DAS_OPT_*)
OPTIONS = {"mc_tma": 0, "mc_tma_ctas": 1, "mc_up_warps": 2, "gemm_2cta": 3, "kc_cluster": 4, "mc_l2_persist": 5}


class DasError(RuntimeError):
    pass


class McDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("T_cap", C.c_int32), ("flags", C.c_int32)]


_vp, _i, _sz, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_h = C.c_void_p      # das_handle*
_PROTOTYPES = {
    "das_strerror": (C.c_char_p, [_i]),
    "das_abi_version": (_i, []),
    "das_last_cuda_error": (_i, []),
    "das_launch_count": (C.c_uint64, []),
    "das_handle_create": (_i, [_i, C.POINTER(_h)]),
    "das_handle_destroy": (_i, [_h]),
    "das_handle_device": (_i, [_h]),
    "das_handle_sm_count": (_i, [_h]),
    "das_handle_l2_info": (_i, [_h, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "das_handle_set_option": (_i, [_h, _i, _i]),
    "das_handle_get_option": (_i, [_h, _i, C.POINTER(_i)]),
    "das_mc_state_bytes": (_i, [C.POINTER(McDesc), C.POINTER(_sz)]),
    "das_mc_reset": (_i, [_h, C.POINTER(McDesc), _vp, _vp]),
    "das_mc_accumulate": (_i, [_h, C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _vp]),
    "das_mc_finalize": (_i, [_h, C.POINTER(McDesc), _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "das_mc_accumulate_finalize": (_i, [_h, C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _vp]),
    "das_mc_upsample_accumulate_finalize": (_i, [_h, C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _i, _vp, _vp,
                                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "das_mc_upsample_supported": (_i, [_h, _i, _i, _i, _i]),
    "das_mc_upsample_variant": (_i, [_h, C.POINTER(McDesc), _i, _i]),
    "das_mc_votes_ptr": (_i, [C.POINTER(McDesc), _vp, C.POINTER(_vp)]),
    "das_suppress_rects": (_i, [_h, _vp, _i, _i, _i, _vp, _i, _vp]),
    "das_suppress_rects_host": (_i, [_h, _vp, _i, _i, _i, C.POINTER(C.c_int32), _i, _vp]),
    "das_add_maps": (_i, [_h, _vp, _vp, _sz, _vp]),
    "das_add_gaussian_noise": (_i, [_h, _vp, _sz, _f, C.c_uint64, C.c_uint64, _vp, _vp]),
    "das_box_sum_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(_sz)]),
    "das_minmax_init": (_i, [_h, _vp, _vp]),
    "das_box_sum": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "das_minmax_normalise": (_i, [_h, _vp, _sz, _vp, _vp]),
    "das_nms_sequences": (_i, [_h, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, C.c_longlong, _vp, _vp]),
    "das_accuracy_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "das_accuracy_scores": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "das_maxsubset_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(_sz)]),
    "das_maxsubset_greedy": (_i, [_h, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "das_topk_workspace_bytes": (_i, [_i, _i, C.POINTER(_sz)]),
    "das_topk": (_i, [_h, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "das_topk_records": (_i, [_h, _vp, _vp, _i, _i, _i, C.c_longlong, _vp, _vp]),
    "das_topk_merge": (_i, [_h, _vp, _i, _i, _i, _vp, _vp]),
    "das_kcenter_filter_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "das_kcenter_filter_build": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _vp]),
    "das_kcenter_filter_stats": (_i, [_h, _vp, _i, _i, _i, C.POINTER(C.c_uint64), _vp]),
    "das_kcenter_init": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "das_kcenter_step": (_i, [_h, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "das_kcenter_workspace_bytes": (_i, [_h, _i, _i, C.POINTER(_sz)]),
    "das_kcenter_greedy": (_i, [_h, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
_handles = {}      # CUDA device ordinal -> das_handle* (one per device for the life of the process)


def load(build_if_missing: bool = True):
    """Load libdas_b200.so (building it in-tree with nvcc if it does not exist yet)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise OSError(f"{LIB_PATH} is missing - run `python -m deep_active_semantic_segmentation_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.das_abi_version() != ABI_VERSION:
        raise DasError(f"libdas_b200 ABI {lib.das_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.das_strerror(status).decode()
        extra = f" (cudaError {lib.das_last_cuda_error()})" if status == -3 else ""
        raise DasError(f"{what or 'libdas_b200'}: {msg}{extra}")


def handle(device=None):
    """The process's das_handle for a CUDA device (ordinal, torch.device or None = torch's current device),
    created on first use.  Options start from the DAS_* environment variables as they are at that moment."""
    import torch

    if device is None:
        idx = torch.cuda.current_device()
    elif isinstance(device, int):
        idx = device
    else:
        device = torch.device(device)
        if device.type != "cuda":
            raise DasError("libdas_b200 works on CUDA devices only (there is no CPU path)")
        idx = torch.cuda.current_device() if device.index is None else device.index
    h = _handles.get(idx)
    if h is None:
        lib = load()
        torch.cuda.init()
        out = _h()
        check(lib.das_handle_create(int(idx), C.byref(out)), "das_handle_create")
        h = _handles[idx] = out
    return h


def set_option(name: str, value: int, device=None) -> int:
    """Set a das_handle option (see OPTIONS) on one device's handle; returns the previous value."""
    lib, h = load(), handle(device)
    old = C.c_int()
    check(lib.das_handle_get_option(h, OPTIONS[name], C.byref(old)), "das_handle_get_option")
    check(lib.das_handle_set_option(h, OPTIONS[name], int(value)), "das_handle_set_option")
    return int(old.value)


def get_option(name: str, device=None) -> int:
    out = C.c_int()
    check(load().das_handle_get_option(handle(device), OPTIONS[name], C.byref(out)), "das_handle_get_option")
    return int(out.value)


def l2_info(device=None) -> dict:
    a, b, c = _sz(), _sz(), _sz()
    check(load().das_handle_l2_info(handle(device), C.byref(a), C.byref(b), C.byref(c)), "das_handle_l2_info")
    return {"l2_bytes": a.value, "persisting_max_bytes": b.value, "window_max_bytes": c.value}


def launch_count() -> int:
    return int(load().das_launch_count())
