"""One-time weight packing: reference state-dict tensors -> the layouts the kernels read.

All of this runs once per ``load_state_dict`` on the parameters' device with torch ops (layout
plumbing; no per-inference arithmetic): eval-mode BatchNorm folded into the adjacent
convolution, convolution weights to K-major [taps*Cin, Cout], LSTM gate rows to slice order.
"""
from __future__ import annotations

import torch

BN_EPS = 1e-5
HU = 8  # hidden units per LSTM CTA slice (csrc/lstm.cu kHU)


def pad_cols(w, mult=4):
    """[K, N] -> [K, ceil(N/mult)*mult] zero padded, contiguous (16-byte aligned rows)."""
    k, n = w.shape
    npad = (n + mult - 1) // mult * mult
    if npad == n:
        return w.contiguous()
    out = w.new_zeros(k, npad)
    out[:, :n] = w
    return out


def bn_fold(gamma, beta, mean, var, eps=BN_EPS):
    """eval BatchNorm as y = x*s + o."""
    s = gamma / torch.sqrt(var + eps)
    return s, beta - mean * s


def pack_conv(w, b, bn):
    """nn.Conv2d weight [Co,Ci,kt,kf] (+bias) followed by eval BN -> (W [kt*kf*Ci, Co], bias [Co]).
    Tap order kt-major then kf (matches the tap tables in the model files)."""
    s, o = bn_fold(*bn)
    wf = w * s[:, None, None, None]
    co, ci, kt, kf = w.shape
    wk = wf.permute(2, 3, 1, 0).reshape(kt * kf * ci, co)
    bias = (b if b is not None else 0.0) * s + o
    return pad_cols(wk), bias.contiguous()


def pack_deconv_parity(w, b, bn):
    """nn.ConvTranspose2d weight [Ci,Co,2,3], stride (1,2), followed by eval BN.
    Returns (W_even [4*Ci, Co], W_odd [2*Ci, Co], bias [Co], fill [Co]) where
      even output columns f'=2m   use taps (kt,kf) in (0,0),(0,2),(1,0),(1,2)
      odd  output columns f'=2m+1 use taps (0,1),(1,1)
    and fill = BN(0) (a zero-padded column that still passes through BN)."""
    if bn is not None:
        s, o = bn_fold(*bn)
    else:
        s, o = torch.ones_like(b), torch.zeros_like(b)
    wf = w * s[None, :, None, None]
    ci, co = w.shape[0], w.shape[1]
    even = torch.stack([wf[:, :, 0, 0], wf[:, :, 0, 2], wf[:, :, 1, 0], wf[:, :, 1, 2]], dim=0)  # [4,Ci,Co]
    odd = torch.stack([wf[:, :, 0, 1], wf[:, :, 1, 1]], dim=0)
    bias = b * s + o
    return (pad_cols(even.reshape(4 * ci, co)), pad_cols(odd.reshape(2 * ci, co)), bias.contiguous(),
            o.contiguous())


DECONV_EVEN_TAPS = [(0, 0), (0, -1), (-1, 0), (-1, -1)]   # (dt, df) for (kt,kf) = (0,0),(0,2),(1,0),(1,2)
DECONV_ODD_TAPS = [(0, 0), (-1, 0)]                        # (kt,kf) = (0,1),(1,1)
CONV23_TAPS = [(kt - 1, kf) for kt in range(2) for kf in range(3)]  # causal k(2,3): in[t-1+kt, 2f+kf]


def slice_rows(hidden, unit_perm=None):
    """Row permutation taking torch's gate-major [i|f|g|o] x H rows to slice order:
    row s*32 + g*8 + j  <-  g*H + unit(s*8+j)."""
    dev = unit_perm.device if unit_perm is not None else None
    units = torch.arange(hidden, device=dev) if unit_perm is None else unit_perm
    s = hidden // HU
    u = units.view(s, 1, HU)                               # [S,1,8]
    g = torch.arange(4, device=units.device).view(1, 4, 1) * hidden
    return (g + u).reshape(-1)                             # [S*4*8]


def split_tf32(x):
    """x -> (hi, lo) with hi = rna_tf32(x), lo = rna_tf32(x - hi): the operand format of the 3xTF32
    tensor-core GEMM (csrc/gemm_tc.cu).  Bit arithmetic = cvt.rna.tf32.f32 (nearest, ties away)."""
    def rna(v):
        return ((v.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = rna(x)
    return hi, rna(x - hi)


def pack_lstm_layer(w_ih, w_hh, b_ih, b_hh, in_perm=None, unit_perm=None, in_scale=None, in_shift=None):
    """Returns a dict: wih_kn [I, 4H] (K-major, for the fp32 FMA GEMM), wih_hi / wih_lo [4H, I]
    (TF32 split, for the tensor-core GEMM), bias [4H], whh [H/8, H, 32], hidden.
    Gate columns / rows are in slice order (see slice_rows).

    in_perm:  new input index -> reference input index (when the producer's layout differs)
    unit_perm: new hidden-unit index -> reference unit index (to emit h in a consumer's layout)
    in_scale/in_shift: an affine on the input (eval BatchNorm1d) folded into Wih / bias.
    """
    hidden = w_hh.shape[1]
    rows = slice_rows(hidden, unit_perm).to(w_ih.device)
    wi = w_ih[rows]
    bias = (b_ih + b_hh)[rows]
    if in_scale is not None:
        bias = bias + wi @ in_shift
        wi = wi * in_scale[None, :]
    if in_perm is not None:
        wi = wi[:, in_perm]
    wh = w_hh[rows]
    if unit_perm is not None:
        wh = wh[:, unit_perm]
    s = hidden // HU
    whp = wh.reshape(s, 4 * HU, hidden).permute(0, 2, 1).contiguous()
    wi = wi.contiguous()
    kin = wi.shape[1]
    kpad = -(-kin // 32) * 32          # the tensor-core GEMM wants K % 32 == 0: zero columns (LSTM.py:17 has K = 161)
    wi_tc = wi if kpad == kin else torch.cat([wi, wi.new_zeros(wi.shape[0], kpad - kin)], dim=1).contiguous()
    hi, lo = split_tf32(wi_tc)
    h16, l16, s16 = pack_linear_f16(wi)
    return {"wih_kn": pad_cols(wi.t().contiguous()), "wih_hi": hi, "wih_lo": lo, "bias": bias.contiguous(),
            "whh": whp, "hidden": hidden, "kin": kin, "wih16_hi": h16, "wih16_lo": l16, "wih16_scale": s16}


def tile_rows(hidden):
    """Row permutation for the fused cell GEMM (csrc/gemm_tc.cu EPI_LSTM_CELL): output column
    j*128 + hf*64 + g*16 + u  <-  torch gate row g*H + 32j + 16hf + u   (u < 16)."""
    j = torch.arange(hidden // 32).view(-1, 1, 1, 1)
    hf = torch.arange(2).view(1, 2, 1, 1)
    g = torch.arange(4).view(1, 1, 4, 1)
    u = torch.arange(16).view(1, 1, 1, 16)
    return (g * hidden + 32 * j + 16 * hf + u).reshape(-1)


def pack_lstm_cell(w_ih, w_hh, b_ih, b_hh, kx_pad=None):
    """[W_ih | W_hh] concatenated along K, rows in tile order, TF32 split.  The input part is zero
    padded to kx_pad columns (a multiple of 32).  Returns dict(w_hi, w_lo, bias, hidden, kx)."""
    hidden = w_hh.shape[1]
    kx = w_ih.shape[1]
    kx_pad = kx_pad or (kx + 31) // 32 * 32
    rows = tile_rows(hidden).to(w_ih.device)
    w = w_ih.new_zeros(4 * hidden, kx_pad + hidden)
    w[:, :kx] = w_ih[rows]
    w[:, kx_pad:] = w_hh[rows]
    hi, lo = split_tf32(w)
    return {"w_hi": hi, "w_lo": lo, "bias": (b_ih + b_hh)[rows].contiguous(), "hidden": hidden, "kx": kx_pad}


F16_ACT_SCALE_LOG2 = 4     # activations of the fp16-pair GEMMs travel as x * 2^4 = hi + lo (|x| < 4094, floor 2e-9)


def split_f16(x, scale_log2=None):
    """x -> (hi, lo, s): IEEE fp16 tensors with x * 2^s = hi + lo, the operand format of the fp16-pair tensor-core GEMM
    (csrc/gemm_tc.cu, kind::f16).  s defaults to the power of two that puts max|x| in [2^13, 2^14) (weights)."""
    x = x.float()
    if scale_log2 is None:
        amax = float(x.abs().max())
        scale_log2 = 0 if amax == 0.0 else 13 - int(torch.floor(torch.log2(torch.tensor(amax, dtype=torch.float64))))
        scale_log2 = max(-14, min(15, scale_log2))
    xs = x * float(2.0 ** scale_log2)
    hi = xs.clamp(-65504.0, 65504.0).to(torch.float16)
    lo = (xs - hi.float()).clamp(-65504.0, 65504.0).to(torch.float16)
    return hi.contiguous(), lo.contiguous(), scale_log2


def pack_lstm_cell_f16(w_ih, w_hh, b_ih, b_hh):
    """pack_lstm_cell for se_lstm_cell_f16x3: [W_ih zero-padded to a multiple of 64 | W_hh], rows in tile order, fp16 pair
    with one per-tensor scale.  Returns dict(w_hi, w_lo, w_scale_log2, bias, hidden, kx)."""
    hidden = w_hh.shape[1]
    kx = w_ih.shape[1]
    kx_pad = (kx + 63) // 64 * 64
    rows = tile_rows(hidden).to(w_ih.device)
    w = w_ih.new_zeros(4 * hidden, kx_pad + hidden)
    w[:, :kx] = w_ih[rows]
    w[:, kx_pad:] = w_hh[rows]
    hi, lo, s = split_f16(w)
    return {"w_hi": hi, "w_lo": lo, "w_scale_log2": s, "bias": (b_ih + b_hh)[rows].contiguous(), "hidden": hidden,
            "kx": kx, "kx_pad": kx_pad}


def pack_conv_f16(w_nk, ntaps, c0, c1):
    """[Cout, ntaps*(c0+c1)] fp32 conv weights (K order: tap, then [source 0 | source 1] channels) -> the fp16-pair
    layout of se_conv_f16x3: every (tap, source) block zero-padded to a multiple of 64 channels.  Returns (hi, lo, s)."""
    cout, k = w_nk.shape
    assert k == ntaps * (c0 + c1)
    p0, p1 = (c0 + 63) // 64 * 64, (c1 + 63) // 64 * 64
    w = w_nk.new_zeros(cout, ntaps, p0 + p1)
    src = w_nk.view(cout, ntaps, c0 + c1)
    w[:, :, :c0] = src[:, :, :c0]
    if c1:
        w[:, :, p0:p0 + c1] = src[:, :, c0:]
    return split_f16(w.view(cout, ntaps * (p0 + p1)))


def pack_linear_f16(w_nk):
    """nn.Linear / LSTM input weights [N, K] -> fp16 pair with K zero-padded to a multiple of 8.  Returns (hi, lo, s)."""
    n, k = w_nk.shape
    kp = (k + 7) // 8 * 8
    if kp != k:
        w_nk = torch.cat([w_nk, w_nk.new_zeros(n, kp - k)], dim=1)
    return split_f16(w_nk.contiguous())
