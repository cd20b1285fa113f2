"""ctypes binding of libdas_b200.so - the only way the Python layer reaches the CUDA kernels.

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised (DasError / OSError).  Prototypes mirror include/das_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdas_b200.so")

ABI_VERSION = 3
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


class DasError(RuntimeError):
    pass


class McDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("T_cap", C.c_int32), ("flags", C.c_int32)]


_vp, _i, _sz, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_PROTOTYPES = {
    "das_strerror": (C.c_char_p, [_i]),
    "das_abi_version": (_i, []),
    "das_last_cuda_error": (_i, []),
    "das_launch_count": (C.c_uint64, []),
    "das_mc_state_bytes": (_i, [C.POINTER(McDesc), C.POINTER(_sz)]),
    "das_mc_reset": (_i, [C.POINTER(McDesc), _vp, _vp]),
    "das_mc_accumulate": (_i, [C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _vp]),
    "das_mc_finalize": (_i, [C.POINTER(McDesc), _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "das_mc_accumulate_finalize": (_i, [C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _vp]),
    "das_mc_upsample_accumulate_finalize": (_i, [C.POINTER(McDesc), _vp, C.POINTER(_vp), _i, _i, _i, _vp, _vp, _vp,
                                                  _vp, _vp, _vp, _vp, _vp, _vp]),
    "das_mc_upsample_supported": (_i, [_i, _i, _i, _i]),
    "das_mc_votes_ptr": (_i, [C.POINTER(McDesc), _vp, C.POINTER(_vp)]),
    "das_suppress_rects": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "das_add_maps": (_i, [_vp, _vp, _sz, _vp]),
    "das_box_sum_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(_sz)]),
    "das_minmax_init": (_i, [_vp, _vp]),
    "das_box_sum": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "das_minmax_normalise": (_i, [_vp, _sz, _vp, _vp]),
    "das_nms_sequences": (_i, [_vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, C.c_longlong, _vp, _vp]),
    "das_accuracy_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "das_accuracy_scores": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "das_maxsubset_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(_sz)]),
    "das_maxsubset_greedy": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "das_topk_workspace_bytes": (_i, [_i, _i, C.POINTER(_sz)]),
    "das_topk": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "das_kcenter_filter_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "das_kcenter_filter_build": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "das_kcenter_filter_stats": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_uint64), _vp]),
    "das_kcenter_init": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "das_kcenter_step": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "das_kcenter_workspace_bytes": (_i, [_i, _i, C.POINTER(_sz)]),
    "das_kcenter_greedy": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


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


def launch_count() -> int:
    return int(load().das_launch_count())
