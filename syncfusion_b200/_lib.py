"""ctypes binding of libsyncfusion_b200.so (include/syncfusion_b200.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ["SFB_LIB"]) if os.environ.get("SFB_LIB") else HERE / "libsyncfusion_b200.so"   # SFB_LIB: A/B a build
SFB_MAX_DEPTH = 16

EXPORTS = [
    "sfb_create", "sfb_destroy", "sfb_last_error", "sfb_set_param", "sfb_finalize", "sfb_workspace_bytes", "sfb_workspace_bytes_m",
    "sfb_unet_forward", "sfb_sample", "sfb_last_launch_count", "sfb_postprocess", "sfb_postprocess_out_len", "sfb_encoder_create", "sfb_encoder_destroy", "sfb_encoder_last_error",
    "sfb_encoder_set_param", "sfb_encoder_finalize", "sfb_encoder_level_length", "sfb_encoder_workspace_bytes", "sfb_encoder_forward", "sfb_dbg_set_op_limit", "sfb_dbg_plan_size", "sfb_dbg_plan_size_m",
    "sfb_dbg_op_info", "sfb_dbg_wait_log", "sfb_dbg_fault_inject", "sfb_dbg_set_grid_limit", "sfb_dbg_sk_timeline", "sfb_dbg_profile", "sfb_dbg_profile_report", "sfb_dbg_gemm", "sfb_dbg_attention",
]


class SfbUnetConfig(C.Structure):
    _fields_ = [
        ("depth", C.c_int32), ("in_channels", C.c_int32),
        ("channels", C.c_int32 * SFB_MAX_DEPTH), ("factors", C.c_int32 * SFB_MAX_DEPTH),
        ("items", C.c_int32 * SFB_MAX_DEPTH), ("attentions", C.c_int32 * SFB_MAX_DEPTH),
        ("cross_attentions", C.c_int32 * SFB_MAX_DEPTH), ("context_channels", C.c_int32 * SFB_MAX_DEPTH),
        ("attention_heads", C.c_int32), ("attention_features", C.c_int32), ("embedding_features", C.c_int32),
        ("embedding_max_length", C.c_int32), ("resnet_groups", C.c_int32), ("modulation_features", C.c_int32),
        ("upsample_mode", C.c_int32), ("precision", C.c_int32),
    ]


class SfbEncoderConfig(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("channels", C.c_int32), ("n_levels", C.c_int32),
        ("multipliers", C.c_int32 * (SFB_MAX_DEPTH + 1)), ("factors", C.c_int32 * SFB_MAX_DEPTH),
        ("num_blocks", C.c_int32 * SFB_MAX_DEPTH), ("resnet_groups", C.c_int32), ("patch_size", C.c_int32),
    ]


class SfbError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) if it is missing or un-loadable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SfbError(f"{LIB_PATH} not found - build it with `python -m syncfusion_b200.build` "
                       "(there is no CPU / PyTorch fallback for the sampling path)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.sfb_create.argtypes = [C.POINTER(SfbUnetConfig), i32, C.POINTER(vp)]
    lib.sfb_destroy.argtypes = [vp]
    lib.sfb_destroy.restype = None
    lib.sfb_last_error.argtypes = [vp]
    lib.sfb_last_error.restype = C.c_char_p
    lib.sfb_set_param.argtypes = [vp, C.c_char_p, vp, i32, C.POINTER(i64), i32]
    lib.sfb_finalize.argtypes = [vp]
    lib.sfb_workspace_bytes.argtypes = [vp, i64, i64, i32, i64, C.POINTER(C.c_size_t)]
    if hasattr(lib, "sfb_workspace_bytes_m"):
        lib.sfb_workspace_bytes_m.argtypes = [vp, i64, i64, i32, i64, i64, C.POINTER(C.c_size_t)]
    lib.sfb_unet_forward.argtypes = [vp, vp, vp, C.POINTER(vp), i32, vp, i64, f32, vp, i64, i64, vp, C.c_size_t, vp]
    lib.sfb_sample.argtypes = [vp, vp, i32, C.POINTER(vp), i32, vp, i64, f32, vp, vp, vp, vp, i64, i64, vp,
                               C.c_size_t, vp]
    lib.sfb_last_launch_count.argtypes = [vp]
    if hasattr(lib, "sfb_encoder_create"):
        lib.sfb_encoder_create.argtypes = [C.POINTER(SfbEncoderConfig), i32, C.POINTER(vp)]
        lib.sfb_encoder_destroy.argtypes = [vp]
        lib.sfb_encoder_destroy.restype = None
        lib.sfb_encoder_last_error.argtypes = [vp]
        lib.sfb_encoder_last_error.restype = C.c_char_p
        lib.sfb_encoder_set_param.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
        lib.sfb_encoder_finalize.argtypes = [vp]
        lib.sfb_encoder_level_length.argtypes = [vp, i64, i32]
        lib.sfb_encoder_level_length.restype = i64
        lib.sfb_encoder_workspace_bytes.argtypes = [vp, i64, i64, C.POINTER(C.c_size_t)]
        lib.sfb_encoder_forward.argtypes = [vp, vp, i64, i64, C.POINTER(vp), i32, vp, C.c_size_t, vp]
    if hasattr(lib, "sfb_postprocess"):
        lib.sfb_postprocess_out_len.argtypes = [i64, i32, i32]
        lib.sfb_postprocess_out_len.restype = i64
        lib.sfb_postprocess.argtypes = [i32, vp, vp, i64, i64, i64, i32, i32, vp, i64, vp, vp]
    lib.sfb_last_launch_count.restype = i64
    lib.sfb_dbg_set_op_limit.argtypes = [vp, i32]
    lib.sfb_dbg_plan_size.argtypes = [vp, i64, i64, i32, vp, C.c_size_t]
    if hasattr(lib, "sfb_dbg_plan_size_m"):
        lib.sfb_dbg_plan_size_m.argtypes = [vp, i64, i64, i32, i64, vp, C.c_size_t]
    lib.sfb_dbg_op_info.argtypes = [vp, i32, C.c_char_p, i32]
    lib.sfb_dbg_sk_timeline.argtypes = [vp, i32, vp, i32]
    if hasattr(lib, "sfb_dbg_set_grid_limit"):
        lib.sfb_dbg_set_grid_limit.argtypes = [vp, i32]
    if not (os.environ.get("SFB_LIB") and not hasattr(lib, "sfb_dbg_wait_log")):   # an older A/B build may lack these two
        lib.sfb_dbg_wait_log.argtypes = [vp, C.c_char_p, i32]
        lib.sfb_dbg_fault_inject.argtypes = [vp, vp]
    lib.sfb_dbg_profile.argtypes = [vp, i32]
    lib.sfb_dbg_profile_report.argtypes = [vp, C.c_char_p, i32]
    lib.sfb_dbg_gemm.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp] + [i32] * 9 + [vp]
    lib.sfb_dbg_attention.argtypes = [i32, vp, vp, i32, i32, vp]
    for name in EXPORTS:          # every symbol the header declares must resolve
        if os.environ.get("SFB_LIB") and (name in ("sfb_dbg_wait_log", "sfb_dbg_fault_inject", "sfb_dbg_set_grid_limit", "sfb_postprocess", "sfb_postprocess_out_len", "sfb_workspace_bytes_m", "sfb_dbg_plan_size_m") or name.startswith("sfb_encoder_")) and not hasattr(lib, name):
            continue
        getattr(lib, name)
    _lib = lib
    return lib
