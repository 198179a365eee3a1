"""One nn.LSTM layer = hoisted input projection (one GEMM over all T) + persistent recurrence.

Input projection: 3xTF32 on tcgen05 (csrc/gemm_tc.cu) whenever there is at least one 128-row tile (an input width that is
not a multiple of 32 is zero-padded while it is split), else the fp32 FMA implicit-GEMM kernel.  Recurrence: csrc/lstm.cu.
Reference: nn.LSTM in CRN/CRN.py:20,29 and LSTM/LSTM.py:17-18,26-27.
"""
from __future__ import annotations

import torch

from . import ops

import os

USE_TENSOR_CORES = True   # flipped by tests / bench A-B runs only
# fp16 operand pairs (tcgen05 kind::f16) instead of TF32 pairs where a model has the f16 packing: the same 22-bit
# products at twice the MMA rate.  SE_F16_PAIRS=0 restores the 3xTF32 path everywhere (A/B runs).
USE_F16_PAIRS = os.environ.get("SE_F16_PAIRS", "1") != "0"


def input_projection(seq2d, layer, pair=None):
    """seq2d [M, K] fp32 (or None when ``pair`` = its TF32 (hi, lo) split is already available)."""
    m, k = (seq2d if seq2d is not None else pair[0]).shape
    n = 4 * layer["hidden"]
    if USE_TENSOR_CORES and USE_F16_PAIRS and "wih16_hi" in layer and (m >= 128 or pair is not None) and \
            (pair is None or pair[0].dtype == torch.float16):
        a = pair if pair is not None else ops.split_f16(seq2d)       # K rounded up to 8, zero tail
        return ops.gemm_f16x3(a, (layer["wih16_hi"], layer["wih16_lo"]), layer["wih16_scale"], layer["bias"], n)
    kw = layer["wih_hi"].shape[1]                # K of the packed tensor-core weights: the input width rounded up to 32
    if USE_TENSOR_CORES and (m >= 128 or pair is not None) and (kw == k or (pair is None and layer.get("kin") == k)):
        if pair is not None:
            a_hi, a_lo = pair
        elif kw == k:
            a_hi, a_lo = ops.split_tf32(seq2d)
        else:                                    # zero-pad the rows on the way to the split (161 -> 192 bins)
            a_hi, a_lo = ops.pad_split_tf32(seq2d, kw)
        return ops.gemm_tf32x3(a_hi, a_lo, layer["wih_hi"], layer["wih_lo"], layer["bias"], n)
    if seq2d is None:
        raise RuntimeError("fp32 activation needed for the FMA projection")
    return ops.linear(seq2d, layer["wih_kn"], layer["bias"], n)


def lstm_layer(seq2d, layer, b, t, pair=None):
    """seq2d [B*T, I] -> hseq [B, T, H]."""
    h = layer["hidden"]
    xp = input_projection(seq2d, layer, pair)
    return ops.lstm_seq(xp.view(b, t, 4 * h), layer["whh"], h)
