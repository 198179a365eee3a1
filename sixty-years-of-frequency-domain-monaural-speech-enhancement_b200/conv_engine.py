"""Convolution dispatch shared by the CRN-family and DCCRN models: every causal Conv2d /
ConvTranspose2d parity class goes either to the tensor-core implicit GEMM (csrc/conv_tc.cu, 3xTF32,
needs channel counts that are multiples of 32) or to the fp32 FMA implicit GEMM (csrc/gemm.cu).

Activations travel between layers as :class:`Act`: the fp32 tensor and/or its (hi, lo) operand pair.
A tensor-core layer writes the pair of its output in its epilogue, so consecutive tensor-core
layers need no separate split pass.  The pair is either a TF32 pair (two fp32 tensors, 3xTF32 kernels)
or an fp16 pair (two fp16 tensors scaled by 2^ops.F16_ACT_SCALE_LOG2, kind::f16 kernels: twice the MMA
rate, half the bytes); a model opts into the latter per activation (``new_act(..., f16=True)``) and
``conv`` follows the dtype of what it is given.
"""
from __future__ import annotations

import os

import torch

from . import lstm_engine, ops, packing


class Act:
    """Channels-last activation [B,T,F,C] as fp32 and/or TF32 split pair."""
    __slots__ = ("f32", "pair")

    def __init__(self, f32=None, pair=None):
        self.f32, self.pair = f32, pair

    @property
    def shape(self):
        return (self.f32 if self.f32 is not None else self.pair[0]).shape

    def get_pair(self, f16=False):
        if self.pair is None:
            if f16:
                x = self.f32
                hi, lo = ops.split_f16(x.view(-1, x.shape[-1]))
                self.pair = (hi.view(x.shape), lo.view(x.shape))
            else:
                self.pair = ops.split_tf32(self.f32)
        return self.pair

    @property
    def is_f16(self):
        return self.pair is not None and self.pair[0].dtype == torch.float16

    def value(self):
        """fp32 value of this activation (taps / tests)."""
        if self.f32 is not None:
            return self.f32
        if self.is_f16:
            return (self.pair[0].float() + self.pair[1].float()) * (2.0 ** -ops.F16_ACT_SCALE_LOG2)
        return self.pair[0] + self.pair[1]

    def get_f32(self):
        if self.f32 is None:
            raise RuntimeError("fp32 copy of this activation was not requested from its producer")
        return self.f32


class ConvWeights:
    """One conv (or one output-column parity of a transposed conv): K-major fp32 for the FMA kernel
    and [Cout, K] TF32 pair for the tensor-core kernel."""
    __slots__ = ("kn", "hi", "lo", "cout", "_f16")

    def __init__(self, w_kn, cout):
        self.kn = packing.pad_cols(w_kn)
        self.hi, self.lo = packing.split_tf32(w_kn[:, :cout].t().contiguous())
        self.cout = cout
        self._f16 = {}

    def f16(self, ntaps, c0, c1, c0p=None, c1p=None):
        """(hi, lo, scale_log2) fp16 pair in the se_conv_f16x3 layout, packed on first use.  c0p / c1p: channel counts the
        ACTIVATIONS are zero-padded to (multiples of 8) when the layer's own are not -- zero weight rows for the padding."""
        c0p, c1p = c0p or c0, c1p or c1
        key = (ntaps, c0, c1, c0p, c1p)
        if key not in self._f16:
            w = self.kn[:, :self.cout]
            if (c0p, c1p) != (c0, c1):
                src = w.view(ntaps, c0 + c1, self.cout)
                wp = w.new_zeros(ntaps, c0p + c1p, self.cout)
                wp[:, :c0] = src[:, :c0]
                if c1:
                    wp[:, c0p:c0p + c1] = src[:, c0:]
                w = wp.view(ntaps * (c0p + c1p), self.cout)
            self._f16[key] = packing.pack_conv_f16(w.t().contiguous(), ntaps, c0p, c1p)
        return self._f16[key]


SMALL_CIN_ON_F16 = os.environ.get("SE_F16_SMALL_CIN", "1") != "0"      # A/B switch for the rule in conv() below


def tc_eligible(c0, c1, cout, fout, sf, f16=False):
    cm = 8 if f16 else 32            # fp16 k-blocks are zero-filled past the last channel
    return (lstm_engine.USE_TENSOR_CORES and c0 % cm == 0 and c1 % cm == 0 and cout % 4 == 0 and cout >= 4
            and fout <= 128 and (fout - 1) * sf + 1 <= 256)


def conv(src: Act, skip, B, T, Fin, Fout, taps, sf, w: ConvWeights, bias, act, dst: Act, dstF, dst_f0=0, dst_fstep=1,
         act_param=0.0, fill_f=-1, fill=None, glu=None):
    """Runs one implicit-GEMM launch writing into ``dst`` (whichever of dst.f32 / dst.pair exist).  ``glu`` = (scale,
    shift): gated conv fused into the tensor-core epilogue (``w`` / ``bias`` with interleaved (conv1, conv2) columns,
    ``dst`` with w.cout / 2 channels); tensor-core layers only."""
    c0 = src.shape[-1]
    c1 = skip.shape[-1] if skip is not None else 0
    f16 = src.is_f16 or (dst.is_f16 and src.pair is None)
    if f16 and tc_eligible(c0, c1, w.cout, Fout, sf, True):
        w_hi, w_lo, ws = w.f16(len(taps), c0, c1)
        ops.conv_f16x3(src.get_pair(True), skip.get_pair(True) if skip is not None else None, B, T, Fin, Fout, taps, sf,
                       w_hi, w_lo, ws, bias, w.cout, act, dstF, dst_f0, dst_fstep, act_param=act_param, out=dst.f32,
                       out_pair16=dst.pair, glu=glu)
        if fill_f >= 0:
            if dst.f32 is not None:
                ops.fill_column(dst.f32, fill, fill_f, act, act_param)
            if dst.pair is not None:
                raise NotImplementedError("fill column on a split-only output")
        return
    if glu is not None and not tc_eligible(c0, c1, w.cout, Fout, sf):
        raise RuntimeError("the fused gate needs a tensor-core eligible layer")
    # Layers the TF32 tensor path cannot take (channel counts that are not multiples of 32: the 2-channel RI inputs of the
    # U2-Net / DCCRN / Uformer encoders, 8 / 16-channel levels) ran on the fp32 FMA implicit GEMM, which is store-bound at
    # wide outputs (TaylorSENet en1, K = 20, N = 128: 2.7 ms per launch).  On fp16 pairs a k-block is zero-filled past the
    # last channel, so they run on the tensor cores after a split of the fp32 activation (channels padded to 8) -- where the
    # output is WIDE (>= 64 channels): TaylorSENet +4 %, CTSNet +7 %, G2Net +1 %.  Narrow outputs stay on the FMA kernel: with
    # 16 / 32 output channels the ten 16-byte-row TMA boxes per tile cost more than the FMA kernel's stores (measured: DCCRN
    # enc0 0.93 ms, Uformer's first level 2.4 ms on the tensor path vs < 0.65 ms on the FMA path; DCCRN / Uformer / DPCRN -3 %).
    c0p, c1p = (c0 + 7) // 8 * 8, (c1 + 7) // 8 * 8
    if (SMALL_CIN_ON_F16 and lstm_engine.USE_F16_PAIRS and glu is None and not tc_eligible(c0, c1, w.cout, Fout, sf)
            and src.f32 is not None and (skip is None or skip.f32 is not None) and w.cout >= 64
            and tc_eligible(c0p, c1p, w.cout, Fout, sf, True) and not dst.is_f16):
        def pair16(a, c, cp):
            x = a.f32
            hi, lo = ops.split_f16(x.reshape(-1, c), kpad=cp)
            return hi.view(*x.shape[:-1], cp), lo.view(*x.shape[:-1], cp)
        w_hi, w_lo, ws = w.f16(len(taps), c0, c1, c0p, c1p)
        ops.conv_f16x3(pair16(src, c0, c0p), pair16(skip, c1, c1p) if skip is not None else None, B, T, Fin, Fout, taps, sf,
                       w_hi, w_lo, ws, bias, w.cout, act, dstF, dst_f0, dst_fstep, act_param=act_param, out=dst.f32,
                       out_pair=dst.pair)
        if fill_f >= 0:
            if dst.f32 is not None:
                ops.fill_column(dst.f32, fill, fill_f, act, act_param)
            if dst.pair is not None:
                raise NotImplementedError("fill column on a split-only output")
        return
    if tc_eligible(c0, c1, w.cout, Fout, sf):
        ops.conv_tf32x3(src.get_pair(), skip.get_pair() if skip is not None else None, B, T, Fin, Fout, taps, sf,
                        w.hi, w.lo, bias, w.cout, act, dstF, dst_f0, dst_fstep, act_param=act_param, out=dst.f32,
                        out_pair=dst.pair, glu=glu)
        if fill_f >= 0:
            if dst.f32 is not None:
                ops.fill_column(dst.f32, fill, fill_f, act, act_param)
            if dst.pair is not None:
                raise NotImplementedError("fill column on a split-only output")
    else:
        if dst.f32 is None:
            raise RuntimeError("fp32 FMA path needs an fp32 destination")
        ops.conv_gemm(src.get_f32(), skip.get_f32() if skip is not None else None, B, T, Fin, Fout, taps, sf, w.kn,
                      bias, w.cout, act, dst.f32, dstF, dst_f0, dst_fstep, fill_f=fill_f, fill=fill,
                      act_param=act_param)
        if dst.pair is not None:
            raise RuntimeError("split output requested from the FMA path: allocate dst without a pair and split")


def merge_parity(w_even: ConvWeights, taps_even, w_odd: ConvWeights, taps_odd):
    """The two output-column parity classes of a stride-2 transposed conv as ONE weight matrix for
    ``conv_parity2``: K order = the even class's taps (the odd class's taps must be among them), columns
    [even class | odd class], zero rows where the odd class does not use a tap."""
    co = w_even.cout
    assert w_odd.cout == co and all(tp in taps_even for tp in taps_odd)
    ct = w_even.kn.shape[0] // len(taps_even)
    assert w_odd.kn.shape[0] == ct * len(taps_odd)
    m = w_even.kn.new_zeros(len(taps_even) * ct, 2 * co)
    m[:, :co] = w_even.kn[:, :co]
    for j, tp in enumerate(taps_odd):
        i = taps_even.index(tp)
        m[i * ct:(i + 1) * ct, co:] = w_odd.kn[j * ct:(j + 1) * ct, :co]
    return ConvWeights(m, 2 * co)


def parity2_eligible(src: Act, skip, w_merged: ConvWeights, fout_even, sf=1):
    """One-launch parity pair: fp16-pair tensor-core layers whose class width fits the epilogue's column groups and is at
    most 64 channels -- wider layers are tensor-bound and the zero taps of the odd class cost more than the second read of
    the activation saves (measured on B200, CRN de1 with 128 channels: 0.38 ms merged vs 0.35 ms as two launches; de2-de4
    with 64 / 32 / 16 channels: 0.66 vs 0.77 ms)."""
    c0 = src.shape[-1]
    c1 = skip.shape[-1] if skip is not None else 0
    co2 = w_merged.cout
    bn = 128 if co2 > 64 else (64 if co2 > 32 else (32 if co2 > 16 else 16))
    return (src.is_f16 or src.pair is None) and tc_eligible(c0, c1, co2, fout_even, sf, True) and \
        co2 % 8 == 0 and (co2 // 2) % (bn // 2) == 0 and co2 // 2 <= 64


def conv_parity2(src: Act, skip, B, T, Fin, fout_even, fout_odd, taps_even, w_merged: ConvWeights, bias, act, dst: Act,
                 dstF, dst_f0=0, act_param=0.0):
    """Both parity classes of a stride-2 transposed conv in one se_conv_f16x3 launch (ncls = 2): even class at output
    columns dst_f0 + 2 m (m < fout_even), odd class at dst_f0 + 1 + 2 m (m < fout_odd)."""
    c0 = src.shape[-1]
    c1 = skip.shape[-1] if skip is not None else 0
    w_hi, w_lo, ws = w_merged.f16(len(taps_even), c0, c1)
    ops.conv_f16x3(src.get_pair(True), skip.get_pair(True) if skip is not None else None, B, T, Fin, fout_even, taps_even,
                   1, w_hi, w_lo, ws, bias, w_merged.cout, act, dstF, dst_f0, 2, act_param=act_param, out=dst.f32,
                   out_pair16=dst.pair, fout1=fout_odd)


def new_act(b, t, f, c, device, want_f32, want_pair, f16=False):
    mk = lambda: torch.empty(b, t, f, c, device=device, dtype=torch.float32)   # noqa: E731
    mk16 = lambda: torch.empty(b, t, f, c, device=device, dtype=torch.float16)   # noqa: E731
    return Act(mk() if want_f32 else None, ((mk16(), mk16()) if f16 else (mk(), mk())) if want_pair else None)
