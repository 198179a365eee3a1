"""One nn.LSTM layer = hoisted input projection (one GEMM over all T) + persistent recurrence.

Input projection: 3xTF32 on tcgen05 (csrc/gemm_tc.cu) whenever the shape allows (K % 32 == 0,
at least one 128-row tile), else the fp32 FMA implicit-GEMM kernel.  Recurrence: csrc/lstm.cu.
Reference: nn.LSTM in CRN/CRN.py:20,29 and LSTM/LSTM.py:17-18,26-27.
"""
from __future__ import annotations

from . import ops

USE_TENSOR_CORES = True   # flipped by tests / bench A-B runs only


def input_projection(seq2d, layer):
    m, k = seq2d.shape
    n = 4 * layer["hidden"]
    if USE_TENSOR_CORES and k % 32 == 0 and m >= 128 and layer["wih_hi"].shape[1] == k:
        a_hi, a_lo = ops.split_tf32(seq2d)
        return ops.gemm_tf32x3(a_hi, a_lo, layer["wih_hi"], layer["wih_lo"], layer["bias"], n)
    return ops.linear(seq2d, layer["wih_kn"], layer["bias"], n)


def lstm_layer(seq2d, layer, b, t):
    """seq2d [B*T, I] -> hseq [B, T, H]."""
    h = layer["hidden"]
    xp = input_projection(seq2d, layer)
    return ops.lstm_seq(xp.view(b, t, 4 * h), layer["whh"], h)
