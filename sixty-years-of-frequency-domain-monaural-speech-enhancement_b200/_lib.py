"""ctypes binding of libse_b200.so (the C ABI declared in include/se_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is
not sm_100 every op raises.  Build with ``python -m build`` in this directory or
``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libse_b200.so")
SE_MAX_TAPS = 16

ACT = {"none": 0, "elu": 1, "softplus": 2, "relu": 3, "sigmoid": 4, "tanh": 5, "prelu": 6}
ISTFT_SPEC, ISTFT_RI_DECOMP, ISTFT_MAG_PHASE, ISTFT_CMASK = 0, 1, 2, 3


class CrnWeights(C.Structure):
    """se_crn_weights (include/se_b200.h): host pointers to the reference state-dict tensors."""
    _fields_ = [
        ("en_w", C.c_void_p * 5), ("en_b", C.c_void_p * 5), ("en_bn", (C.c_void_p * 4) * 5),
        ("lstm_w_ih", C.c_void_p * 2), ("lstm_w_hh", C.c_void_p * 2), ("lstm_b_ih", C.c_void_p * 2),
        ("lstm_b_hh", C.c_void_p * 2),
        ("de_w", C.c_void_p * 5), ("de_b", C.c_void_p * 5), ("de_bn", (C.c_void_p * 4) * 5),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("C0", C.c_int), ("C1", C.c_int),
        ("B", C.c_int), ("T", C.c_int), ("Fin", C.c_int), ("Fout", C.c_int), ("ntaps", C.c_int),
        ("dt", C.c_int * SE_MAX_TAPS), ("df", C.c_int * SE_MAX_TAPS), ("sf", C.c_int),
        ("W", C.c_void_p), ("ldw", C.c_int), ("bias", C.c_void_p), ("Cout", C.c_int), ("act", C.c_int), ("act_param", C.c_float),
        ("dst", C.c_void_p), ("dstF", C.c_int), ("dst_f0", C.c_int), ("dst_fstep", C.c_int),
        ("fill_f", C.c_int), ("fill", C.c_void_p),
    ]


class ConvTcDesc(C.Structure):
    _fields_ = [
        ("src0_hi", C.c_void_p), ("src0_lo", C.c_void_p), ("src1_hi", C.c_void_p), ("src1_lo", C.c_void_p),
        ("C0", C.c_int), ("C1", C.c_int), ("B", C.c_int), ("T", C.c_int), ("Fin", C.c_int), ("Fout", C.c_int),
        ("ntaps", C.c_int), ("dt", C.c_int * SE_MAX_TAPS), ("df", C.c_int * SE_MAX_TAPS), ("sf", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("bias", C.c_void_p), ("Cout", C.c_int), ("act", C.c_int),
        ("act_param", C.c_float), ("out", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("dstF", C.c_int), ("dst_f0", C.c_int), ("dst_fstep", C.c_int),
        ("glu", C.c_int), ("glu_scale", C.c_void_p), ("glu_shift", C.c_void_p),
    ]


class ConvF16Desc(C.Structure):
    """se_conv_f16_desc (include/se_b200.h)."""
    _fields_ = [
        ("src0_hi", C.c_void_p), ("src0_lo", C.c_void_p), ("src1_hi", C.c_void_p), ("src1_lo", C.c_void_p),
        ("C0", C.c_int), ("C1", C.c_int), ("B", C.c_int), ("T", C.c_int), ("Fin", C.c_int), ("Fout", C.c_int),
        ("ntaps", C.c_int), ("dt", C.c_int * SE_MAX_TAPS), ("df", C.c_int * SE_MAX_TAPS), ("sf", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("scale_log2_a", C.c_int), ("scale_log2_w", C.c_int),
        ("bias", C.c_void_p), ("Cout", C.c_int), ("act", C.c_int), ("act_param", C.c_float),
        ("out", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("out16_hi", C.c_void_p), ("out16_lo", C.c_void_p), ("out16_scale_log2", C.c_int),
        ("dstF", C.c_int), ("dst_f0", C.c_int), ("dst_fstep", C.c_int),
        ("glu", C.c_int), ("glu_scale", C.c_void_p), ("glu_shift", C.c_void_p),
        ("ncls", C.c_int), ("fout1", C.c_int),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/se_b200.h declares
_LL, _I, _F, _P = C.c_longlong, C.c_int, C.c_float, C.c_void_p
PROTOTYPES = {
    "se_abi_version": (_I, []),
    "se_last_error": (C.c_char_p, []),
    "se_device_check": (_I, []),
    "se_launch_count": (C.c_ulonglong, []),
    "se_rms_scale": (_I, [_P, _LL, _I, _I, _I, _P, _P, _P]),
    "se_rms_scale_len": (_I, [_P, _LL, _I, _I, _P, _I, _P, _P, _P]),
    "se_stft": (_I, [_P, _LL, _I, _I, _P, _I, _I, _I, _I, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _LL, _F, _F, _P]),
    "se_stft_len": (_I, [_P, _LL, _I, _I, _P, _P, _I, _I, _I, _I, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _LL, _F, _F, _P]),
    "se_istft_len": (_I, [_I, _P, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _LL, _F, _F, _I, _I, _I, _I, _I, _P, _P, _LL,
                          _I, _P, _P]),
    "se_istft": (_I, [_I, _P, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _LL, _F, _F, _I, _I, _I, _I, _I, _P, _P, _LL,
                      _I, _P]),
    "se_conv_gemm": (_I, [C.POINTER(ConvDesc), _P]),
    "se_fill_column": (_I, [_P, _LL, _I, _I, _I, _P, _I, _F, _P]),
    "se_conv_in1": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P]),
    "se_deconv_out1": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _F, _I, _P, _P]),
    "se_lstm_seq": (_I, [_P, _LL, _P, _I, _I, _I, _P, _LL, _LL, _P, _P, _P]),
    "se_lstm_seq_multi": (_I, [_P, _LL, _LL, _P, _LL, _I, _I, _I, _I, _P, _LL, _LL, _LL, _P, _P, _P]),
    "se_set_lstm_engine": (_I, [_I]),
    "se_set_gemm_engine": (_I, [_I]),
    "se_lstm_seq_work_bytes": (_LL, [_I, _I]),
    "se_split_tf32": (_I, [_P, _P, _P, _LL, _P]),
    "se_pad_split_tf32": (_I, [_P, _LL, _I, _I, _P, _P, _P]),
    "se_gemm_tf32x3": (_I, [_P, _P, _LL, _P, _P, _LL, _I, _I, _I, _P, _I, _P, _LL, _P]),
    "se_gemm_tf32x3_ex": (_I, [_P, _P, _LL, _P, _P, _LL, _I, _I, _I, _P, _I, _F, _F, _P, _P, _P, _P, _LL, _P]),
    "se_lstm_cell_tf32x3": (_I, [_P, _P, _LL, _I, _P, _P, _LL, _I, _P, _P, _LL, _P, _I, _P, _P, _P, _P, _P]),
    "se_lstm_cell_tf32x3_ex": (_I, [_P, _P, _LL, _I, _P, _P, _LL, _I, _P, _P, _LL, _P, _I, _P, _P, _P, _P, _LL, _I, _P]),
    "se_fsn_clip_inv_mean": (_I, [_P, _LL, _LL, _LL, _I, _I, _I, _P, _P, _LL, _LL, C.c_double, _P, _P]),
    "se_fsn_fb_input": (_I, [_P, _LL, _LL, _LL, _I, _I, _I, _I, _P, _P, _P, _P]),
    "se_fsn_sb_assemble": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "se_fsn_sb_fc": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "se_fsn_sb_assemble_f16": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "se_split_f16": (_I, [_P, _LL, _I, _LL, _I, _I, _P, _P, _P]),
    "se_gemm_f16x3": (_I, [_P, _P, _LL, _P, _P, _LL, _I, _I, _I, _I, _P, _I, _F, _F, _P, _P, _P, _P, _P, _P, _I, _LL, _P]),
    "se_lstm_cell_f16x3": (_I, [_P, _P, _LL, _I, _P, _P, _LL, _I, _P, _P, _LL, _I, _I, _P, _I, _P, _P, _P, _P, _LL, _I, _P]),
    "se_conv_tf32x3": (_I, [C.POINTER(ConvTcDesc), _P]),
    "se_conv_f16x3": (_I, [C.POINTER(ConvF16Desc), _P]),
    "se_uf_prep": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "se_uf_fusion": (_I, [_P, _P, _LL, _I, _P, _P, _P]),
    "se_uf_fusion_ex": (_I, [_P, _P, _LL, _I, _P, _P, _P, _P, _P, _P, _P]),
    "se_group_layernorm": (_I, [_P, _P, _LL, _I, _I, _P, _P, _F, _I, _F, _P, _P, _P, _P, _P, _P]),
    "se_attention": (_I, [_P, _I, _I, _P, _P, _I, _I, _LL, _I, _LL, _I, _LL, _F, _P, _I, _P]),
    "se_uf_mask": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "se_debug_lstm_tc_profile": (_I, [_P, _I, _I]),
    "se_glu_affine_act": (_I, [_P, _LL, _I, _P, _P, _I, _F, _P, _P, _P, _P]),
    "se_unary": (_I, [_P, _LL, _I, _F, _P, _P, _P, _P]),
    "se_cmul": (_I, [_P, _P, _LL, _P, _P]),
    "se_dccrn_mask": (_I, [_P, _P, _P, _LL, _LL, _LL, _I, _I, _I, _P, _P, _LL, _LL, _LL, _P]),
    "se_dccrn_mask_ex": (_I, [_P, _P, _P, _LL, _LL, _LL, _I, _I, _I, _I, _P, _P, _LL, _LL, _LL, _P]),
    "se_resample": (_I, [_P, _LL, _I, _I, _P, _LL, _I, _I, C.c_double, _P, _P, _P, _I, _I, _P]),
    "se_chan_stats_ws_bytes": (_LL, [_I, _LL, _I]),
    "se_chan_stats": (_I, [_P, _I, _LL, _I, _I, _I, _P, _F, _P, _P, _P, _P]),
    "se_cum_stats": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _F, _P, _P, _P, _P]),
    "se_chan_norm": (_I, [_P, _I, _LL, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P]),
    "se_add": (_I, [_P, _P, _LL, _P, _P, _P, _P]),
    "se_axpby": (_I, [_P, _P, _F, _F, _LL, _P, _P, _P, _P]),
    "se_gaf_update": (_I, [_P, _P, _LL, _I, _P, _P, _LL, _I, _I, _I, _P, _P, _P, _P]),
    "se_taylor_zero": (_I, [_P, _P, _LL, _I, _I, _P, _P, _P, _P]),
    "se_cts_glue1": (_I, [_P, _P, _LL, _P, _P]),
    "se_cts_glue2": (_I, [_P, _P, _P, _LL, _P, _P]),
    "se_plan_create_crn": (_I, [C.POINTER(CrnWeights), _I, _I, C.POINTER(C.c_void_p)]),
    "se_query_workspace": (C.c_longlong, [_P]),
    "se_plan_set_graph": (_I, [_P, _I]),
    "se_forward_crn": (_I, [_P, _P, _P, _I, _I, _P]),
    "se_enhance_crn": (_I, [_P, _P, _LL, _P, _LL, _I, _I, _P, _F, _P]),
    "se_plan_destroy": (_I, [_P]),
}
NORM_PRE = {"none": 0, "glu": 1, "prelu": 2, "glu_prelu": 3}
NORM_POST = {"none": 0, "prelu": 1, "fir": 2}

_lib = None


class SeB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeB200Error(
            f"{LIB_PATH} not found: the CUDA library has not been built (run __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.se_abi_version() != 1:
        raise SeB200Error("ABI version mismatch between _lib.py and libse_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().se_last_error().decode(errors="replace")
        raise SeB200Error(f"{what} failed with status {rc}: {msg}")
