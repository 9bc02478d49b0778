"""ctypes binding of the C-ABI library (include/wxformer_b200.h).

There is no fallback: if ``libwxformer_b200.so`` is missing or a call fails, a ``RuntimeError`` is raised.
"""

from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "csrc", "libwxformer_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "wxformer_b200.h")

WXF_ABI_VERSION = 17

PAD_EARTH, PAD_MIRROR = 0, 1
ACT_NONE, ACT_GELU = 0, 1
ATTN_SHORT, ATTN_LONG = 0, 1


class WxfConvDesc(Structure):
    _fields_ = [
        ("inp", c_void_p), ("w", c_void_p), ("taps", c_void_p), ("bias", c_void_p), ("res", c_void_p),
        ("out", c_void_p),
        ("B", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("lda", c_int32), ("Cin", c_int32),
        ("N", c_int32), ("T", c_int32), ("stride", c_int32),
        ("Ho", c_int32), ("Wo", c_int32),
        ("phases", c_int32), ("out_scale", c_int32),
        ("ldc", c_int32), ("c_off", c_int32),
        ("ldr", c_int32), ("r_off", c_int32),
        ("act", c_int32), ("bias_phase_stride", c_int32),
    ]


class WxfGemmDesc(Structure):
    _fields_ = [
        ("a_hi", c_void_p), ("a_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p),
        ("bias", c_void_p), ("res", c_void_p), ("out", c_void_p), ("out_hi", c_void_p), ("out_lo", c_void_p),
        ("M", c_int64),
        ("N", c_int32), ("K", c_int32), ("lda", c_int32),
        ("ldc", c_int32), ("c_off", c_int32), ("ldr", c_int32), ("r_off", c_int32), ("ldh", c_int32),
        ("act", c_int32), ("w_scale_log2", c_int32),
    ]


class WxfWaterDesc(Structure):
    _fields_ = [
        ("q_pred", c_void_p), ("q_pred_bs", c_int64), ("q_pred_ls", c_int64),
        ("sp_pred", c_void_p), ("sp_pred_bs", c_int64),
        ("q_in", c_void_p), ("q_in_bs", c_int64), ("q_in_ls", c_int64),
        ("sp_in", c_void_p), ("sp_in_bs", c_int64),
        ("precip", c_void_p), ("precip_bs", c_int64),
        ("evapor", c_void_p), ("evapor_bs", c_int64),
        ("area", c_void_p), ("coef_a", c_void_p), ("coef_b", c_void_p),
        ("p0", c_int64), ("np", c_int64),
        ("B", c_int32), ("L", c_int32),
        ("n_seconds", c_float),
    ]


class WxfEnergyDesc(Structure):
    _fields_ = [
        ("t_pred", c_void_p), ("q_pred", c_void_p), ("u_pred", c_void_p), ("v_pred", c_void_p),
        ("pred3_bs", c_int64), ("pred3_ls", c_int64),
        ("sp_pred", c_void_p), ("toa_up_solar", c_void_p), ("toa_up_olr", c_void_p), ("surf_down_solar", c_void_p),
        ("surf_up_solar", c_void_p), ("surf_down_lw", c_void_p), ("surf_up_lw", c_void_p), ("surf_sh", c_void_p),
        ("surf_lh", c_void_p),
        ("pred2_bs", c_int64),
        ("t_in", c_void_p), ("q_in", c_void_p), ("u_in", c_void_p), ("v_in", c_void_p),
        ("in3_bs", c_int64), ("in3_ls", c_int64),
        ("sp_in", c_void_p), ("sp_in_bs", c_int64),
        ("toa_down_in", c_void_p), ("toa_down_bs", c_int64),
        ("gph_surf", c_void_p), ("area", c_void_p), ("coef_a", c_void_p), ("coef_b", c_void_p),
        ("p0", c_int64), ("np", c_int64),
        ("B", c_int32), ("L", c_int32),
        ("n_seconds", c_float),
    ]


class WxfConvTcDesc(Structure):
    _fields_ = [
        ("in_hi", c_void_p), ("in_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p), ("taps", c_void_p),
        ("bias", c_void_p), ("res", c_void_p), ("out", c_void_p), ("out_hi", c_void_p), ("out_lo", c_void_p),
        ("B", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("lda", c_int32), ("Cin", c_int32), ("cin_pad", c_int32),
        ("N", c_int32), ("T", c_int32), ("stride", c_int32),
        ("Ho", c_int32), ("Wo", c_int32),
        ("phases", c_int32), ("out_scale", c_int32),
        ("ldc", c_int32), ("c_off", c_int32), ("ldr", c_int32), ("r_off", c_int32), ("ldh", c_int32), ("h_off", c_int32),
        ("act", c_int32), ("w_scale_log2", c_int32), ("bias_phase_stride", c_int32),
    ]


class WxfToeplitzDesc(Structure):
    _fields_ = [
        ("in_hi", c_void_p), ("in_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p), ("bias", c_void_p),
        ("out", c_void_p),
        ("B", c_int32), ("Hi", c_int32), ("Wi", c_int32), ("lda", c_int32), ("Cin", c_int32), ("cin_pad", c_int32),
        ("ch", c_int32), ("kernel", c_int32), ("pad", c_int32),
        ("Ho", c_int32), ("Wo", c_int32),
        ("ldc", c_int32), ("c_off", c_int32),
        ("w_scale_log2", c_int32), ("oy_off", c_int32),
    ]


_SIGNATURES = {
    "wxf_abi_version": (c_int, []),
    "wxf_last_error": (c_char_p, []),
    "wxf_pad_to_pixel_major": (c_int, [c_void_p, c_void_p] + [c_int] * 13 + [c_void_p]),
    "wxf_pad_to_pixel_major_f16x2": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 13 + [c_void_p]),
    "wxf_preblock_pad_to_pixel_major": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 13
                                        + [c_void_p]),
    "wxf_cross_embed_toeplitz_tc": (c_int, [POINTER(WxfToeplitzDesc), c_void_p]),
    "wxf_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "wxf_layernorm_f16x2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_float,
                                    c_void_p]),
    "wxf_conv_igemm_f32": (c_int, [POINTER(WxfConvDesc), c_void_p]),
    "wxf_gemm_f16x2_tc": (c_int, [POINTER(WxfGemmDesc), c_void_p]),
    "wxf_conv_f16x2_tc": (c_int, [POINTER(WxfConvTcDesc), c_void_p]),
    "wxf_groupnorm_silu_f16x2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    "wxf_split_f16x2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "wxf_window_attention_f16x2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int] + [c_int] * 7
                                   + [c_float, c_void_p]),
    "wxf_window_attention_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int] + [c_int] * 7 + [c_float, c_void_p]),
    "wxf_attention_bias_tile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "wxf_window_attention_tc": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int] + [c_int] * 7
                                + [c_float, c_void_p]),
    "wxf_groupnorm_scratch_bytes": (c_int64, [c_int, c_int64, c_int]),
    "wxf_groupnorm_stats": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_void_p]),
    "wxf_groupnorm_silu": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                   c_int, c_int64, c_int, c_int, c_void_p]),
    "wxf_groupnorm_sums": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "wxf_groupnorm_stats_from_sums": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_double, c_float, c_void_p]),
    "wxf_gather_rows": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "wxf_unpad_resize_to_nchw": (c_int, [c_void_p, c_int, c_void_p] + [c_int] * 12 + [c_void_p]),
    "wxf_layernorm_residual": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                       c_void_p, c_int64, c_int, c_float, c_void_p]),
    "wxf_swin_window_attention": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
                                  + [c_int] * 10 + [c_void_p]),
    "wxf_gather_rows_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int64,
                                   c_int, c_void_p]),
    "wxf_unpatchify_unpad_resize_to_nchw": (c_int, [c_void_p, c_void_p] + [c_int] * 16 + [c_void_p]),
    "wxf_peer_alloc": (c_int, [POINTER(c_void_p), c_int64]),
    "wxf_peer_free": (c_int, [c_void_p]),
    "wxf_peer_export": (c_int, [c_void_p, c_void_p]),
    "wxf_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "wxf_peer_close": (c_int, [c_void_p]),
    "wxf_peer_epoch_advance": (c_int, [c_void_p, c_void_p]),
    "wxf_peer_put": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int, POINTER(c_void_p), c_int, c_void_p,
                             c_void_p]),
    "wxf_peer_scatter_rows": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_void_p), c_int,
                                      c_int, c_int64, c_int, c_void_p, c_void_p]),
    "wxf_peer_wait": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_void_p]),
    "wxf_sum_rank_slots": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "wxf_history_update": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_int64, c_void_p]),
    "wxf_unpad_resize_post_to_nchw": (c_int, [c_void_p, c_int, c_void_p] + [c_int] * 12 + [c_void_p] * 5),
    "wxf_dry_mass_scratch_bytes": (c_int64, [c_int]),
    "wxf_dry_mass_sums": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "wxf_scale_planes": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "wxf_budget_scratch_bytes": (c_int64, [c_int]),
    "wxf_water_budget_sums": (c_int, [POINTER(WxfWaterDesc), c_void_p, c_void_p, c_void_p]),
    "wxf_energy_budget_sums": (c_int, [POINTER(WxfEnergyDesc), c_void_p, c_void_p, c_void_p]),
    "wxf_energy_fix_temperature": (c_int, [POINTER(WxfEnergyDesc), c_void_p, c_void_p]),
    "wxf_noise_coef": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, ctypes.c_uint64, c_void_p, c_int, c_void_p]),
    "wxf_noise_inject": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                 c_int64, c_int, ctypes.c_uint64, c_void_p, c_int, c_void_p]),
    "wxf_noise_step_advance": (c_int, [c_void_p, c_void_p]),
    "wxf_copy_channels": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int64, POINTER(c_int32), POINTER(c_int32),
                                  POINTER(c_int32), c_int, c_void_p]),
}

_lib = None


def declared_symbols(header_path: str = HEADER_PATH):
    """Every function name the public header declares (used by the CPU export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wxf_[a-z0-9_]+)\s*\(", text)))


def load(path: str = LIB_PATH):
    """Load the shared library; raises if it is not built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(path):
        raise RuntimeError(
            f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = int(lib.wxf_abi_version())
    if got != WXF_ABI_VERSION:  # a stale .so would be called with mismatched descriptor layouts
        raise RuntimeError(f"{path} has ABI version {got}, this package needs {WXF_ABI_VERSION}: rebuild it "
                           "(python -c 'import __graft_entry__ as g; g.build()')")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().wxf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")
