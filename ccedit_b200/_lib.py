"""ctypes binding of the C-ABI shared library (include/ccedit_b200.h).

The library is built in-tree by ``ccedit_b200.build`` (nvcc, sm_100a) into ``ccedit_b200/lib/libccedit_b200.so``.
There is no CPU fallback: importing works anywhere (so host logic can be tested on a CPU box), but any compute call
without the library, or without a CUDA device, raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libccedit_b200.so")

MAX_TAPS = 9
GEMM_SILU = 1
GEMM_GEGLU = 2


class GemmDesc(C.Structure):
    """Mirror of ``ccedit_gemm_desc`` (include/ccedit_b200.h)."""

    _fields_ = [
        ("a", C.c_void_p),
        ("a_dims", C.c_int32 * 5),
        ("a_strides", C.c_int64 * 4),
        ("box", C.c_int32 * 4),
        ("out_dims", C.c_int32 * 4),
        ("ntaps", C.c_int32),
        ("taps", (C.c_int32 * 4) * MAX_TAPS),
        ("w", C.c_void_p),
        ("n", C.c_int32),
        ("kpad", C.c_int32),
        ("bn", C.c_int32),
        ("out", C.c_void_p),
        ("out_strides", C.c_int64 * 4),
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p),
        ("rb_dim", C.c_int32),
        ("rb_div", C.c_int32),
        ("rb_ld", C.c_int32),
        ("res1", C.c_void_p),
        ("res1_strides", C.c_int64 * 4),
        ("res2", C.c_void_p),
        ("res2_strides", C.c_int64 * 4),
        ("flags", C.c_int32),
        ("rowstats", C.c_void_p),
        ("colsum", C.c_void_p),
        ("stats_out", C.c_void_p),
        ("rowstats_slots", C.c_int32),
        ("ln_eps", C.c_float),
    ]


class AttnDesc(C.Structure):
    """Mirror of ``ccedit_attn_desc``."""

    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64), ("q_frame_stride", C.c_int64),
        ("o", C.c_void_p), ("ldo", C.c_int64), ("o_frame_stride", C.c_int64),
        ("nseg", C.c_int32),
        ("k", C.c_void_p * 2), ("v", C.c_void_p * 2),
        ("ldk", C.c_int64 * 2), ("ldv", C.c_int64 * 2), ("kv_frame_stride", C.c_int64 * 2),
        ("lkv", C.c_int32 * 2), ("kv_div", C.c_int32 * 2), ("kv_mul", C.c_int32 * 2), ("kv_add", C.c_int32 * 2),
        ("frames", C.c_int32), ("lq", C.c_int32), ("heads", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/ccedit_b200.h declares
SIGNATURES = {
    "ccedit_last_error": (C.c_char_p, []),
    "ccedit_abi_version": (C.c_int, []),
    "ccedit_launch_count": (C.c_int64, []),
    "ccedit_gemm": (C.c_int, [C.POINTER(GemmDesc), _vp]),
    "ccedit_gemm_trace": (C.c_int, [_vp]),
    "ccedit_groupnorm_spatial": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _vp]),
    "ccedit_groupnorm_temporal": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "ccedit_layernorm": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "ccedit_layernorm_stats": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _f32, _vp]),
    "ccedit_layernorm_stats_combine": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _f32, _vp]),
    "ccedit_attention": (C.c_int, [C.POINTER(AttnDesc), _vp]),
    "ccedit_temporal_attention": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32,
                                            _f32, _vp]),
    "ccedit_ncthw_to_cl": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp]),
    "ccedit_out_temporal": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_timestep_embedding": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _vp]),
    "ccedit_linear_small": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_parity_split": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_upsample_nearest2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_add_rows": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "ccedit_add_center_frame": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_to_half": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ccedit_hint_stem01": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ccedit_hint_stem23": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ccedit_softmax_rows": (C.c_int, [_vp, _i64, _i64, _i32, _vp]),
    "ccedit_cl_to_ncthw": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _vp]),
    "ccedit_embed_tokens": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccedit_quick_gelu": (C.c_int, [_vp, _i64, _vp]),
    "ccedit_causal_attention_small": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp]),
    "ccedit_sampler_prepare": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "ccedit_sampler_mid": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp]),
    "ccedit_sampler_final": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i64, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"ccedit_b200: native library {LIB_PATH} is missing - run `python -m ccedit_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback for the kernels"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().ccedit_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"ccedit_b200 native call failed ({what}): {msg} [status {status}]")


def launch_count() -> int:
    return int(load().ccedit_launch_count())
