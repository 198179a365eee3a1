"""Thin torch-tensor wrappers over the C ABI (device pointers + current stream).  Plumbing only:
no arithmetic happens in Python."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import ACT, ConvDesc, ConvF16Desc, ConvTcDesc, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise _lib.SeB200Error("se_b200 ops need CUDA float32 tensors (no CPU fallback)")


_checked = False

# ---- optional per-op CUDA-event timing (bench.py's roofline leg) ---------------------------------
_recorder = None   # dict: op name -> list of (start_event, end_event)


class _Timed:
    __slots__ = ("name", "s")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _recorder is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if _recorder is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            _recorder.setdefault(self.name, []).append((self.s, e))
        return False


def start_recording():
    global _recorder
    _recorder = {}


def stop_recording():
    """Returns {op: (launch_count, total_ms)}; call after torch.cuda.synchronize()."""
    global _recorder
    rec, _recorder = _recorder, None
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (rec or {}).items()}


def device_check():
    global _checked
    if not _checked:
        check(_lib.load().se_device_check(), "se_device_check")
        _checked = True


def launch_count() -> int:
    return int(_lib.load().se_launch_count())


def _check_lengths(lengths, b, device):
    if lengths is None:
        return
    if not (lengths.is_cuda and lengths.dtype == torch.int32 and lengths.is_contiguous() and lengths.numel() == b):
        raise _lib.SeB200Error("lengths must be a contiguous CUDA int32 tensor with one entry per clip")


def rms_scale(wav, reciprocal=False, lengths=None):
    """wav [B,N] -> (c [B], inv_c [B]).  a1.  ``lengths`` (int32 [B], CUDA): per-clip sample counts of a tail-padded
    batch (see se_rms_scale_len)."""
    _need_cuda(wav)
    device_check()
    assert wav.dim() == 2 and wav.stride(1) == 1
    b, n = wav.shape
    _check_lengths(lengths, b, wav.device)
    c = torch.empty(b, device=wav.device, dtype=torch.float32)
    ic = torch.empty_like(c)
    with _Timed("rms_scale"):
        check(_lib.load().se_rms_scale_len(_ptr(wav), wav.stride(0), b, n, _ptr(lengths), int(reciprocal), _ptr(c),
                                           _ptr(ic), _stream()), "se_rms_scale")
    return c, ic


def _plane_strides(t, layout):
    """Strides (sb, st, sf) of a [B,T,F] ('btf') or [B,F,T] ('bft') plane view."""
    if layout == "btf":
        return t.stride(0), t.stride(1), t.stride(2)
    return t.stride(0), t.stride(2), t.stride(1)


def stft(wav, scale, n_fft, win, hop, mag=None, re=None, im=None, layout="btf", p_mag=1.0, p_ri=1.0, lengths=None):
    """Fused STFT.  mag / re / im are pre-allocated planes ([B,T,F] for 'btf', [B,F,T] for 'bft');
    re and im must share strides, mag has its own.  ``lengths``: see se_stft_len (tail-padded batch)."""
    _need_cuda(wav, scale, mag, re, im)
    device_check()
    b, n = wav.shape
    _check_lengths(lengths, b, wav.device)
    t = 1 + n // hop
    msb = mst = msf = sb = st = sf = 0
    if mag is not None:
        msb, mst, msf = _plane_strides(mag, layout)
    if re is not None:
        sb, st, sf = _plane_strides(re, layout)
        assert _plane_strides(im, layout) == (sb, st, sf), "re and im planes must share strides"
    with _Timed("stft"):
        check(_lib.load().se_stft_len(_ptr(wav), wav.stride(0), b, n, _ptr(lengths), _ptr(scale), n_fft, win, hop, t,
                                      _ptr(mag), msb, mst, msf, _ptr(re), _ptr(im), sb, st, sf, float(p_mag),
                                      float(p_ri), _stream()), "se_stft")
    return t


def istft(mode, a_re, a_im, b_re, b_im, n_fft, win, hop, out, length, out_scale=None, inv_p=1.0, p_x=1.0,
          layout_a="btf", layout_b="btf", lengths=None):
    _need_cuda(a_re, a_im, b_re, b_im, out, out_scale)
    device_check()
    _check_lengths(lengths, a_re.shape[0], a_re.device)
    asb, ast, asf = _plane_strides(a_re, layout_a)
    if a_im is not None:
        assert _plane_strides(a_im, layout_a) == (asb, ast, asf)
    if b_re is not None:
        bsb, bst, bsf = _plane_strides(b_re, layout_b)
        assert _plane_strides(b_im, layout_b) == (bsb, bst, bsf)
    else:
        bsb = bst = bsf = 0
    bsz = a_re.shape[0]
    t = a_re.shape[1] if layout_a == "btf" else a_re.shape[2]
    with _Timed("istft"):
        check(_lib.load().se_istft_len(mode, _ptr(a_re), _ptr(a_im), asb, ast, asf, _ptr(b_re), _ptr(b_im), bsb, bst,
                                       bsf, float(inv_p), float(p_x), bsz, t, n_fft, win, hop, _ptr(out_scale),
                                       _ptr(out), out.stride(0), int(length), _ptr(lengths), _stream()), "se_istft")
    return out


def conv_gemm(src0, src1, B, T, Fin, Fout, taps, sf, W, bias, Cout, act, dst, dstF, dst_f0=0, dst_fstep=1,
              fill_f=-1, fill=None, act_param=0.0):
    """Implicit-GEMM conv.  taps: list of (dt, df).  W [ntaps*(C0+C1), ldw] fp32 (K-major)."""
    _need_cuda(src0, src1, W, bias, dst, fill)
    device_check()
    d = ConvDesc()
    d.src0, d.src1 = src0.data_ptr(), (src1.data_ptr() if src1 is not None else 0)
    c0 = src0.shape[-1]
    c1 = src1.shape[-1] if src1 is not None else 0
    d.C0, d.C1 = c0, c1
    d.B, d.T, d.Fin, d.Fout = B, T, Fin, Fout
    d.ntaps = len(taps)
    for i, (dt, df) in enumerate(taps):
        d.dt[i], d.df[i] = dt, df
    d.sf = sf
    assert W.is_contiguous() and W.shape[0] == len(taps) * (c0 + c1), (W.shape, len(taps), c0, c1)
    d.W, d.ldw = W.data_ptr(), W.shape[1]
    d.bias = bias.data_ptr() if bias is not None else 0
    d.Cout, d.act, d.act_param = Cout, ACT[act], float(act_param)
    d.dst, d.dstF, d.dst_f0, d.dst_fstep = dst.data_ptr(), dstF, dst_f0, dst_fstep
    d.fill_f = fill_f
    d.fill = fill.data_ptr() if fill is not None else 0
    with _Timed(f"conv_gemm[K={len(taps) * (c0 + c1)},N={Cout}]"):
        check(_lib.load().se_conv_gemm(C.byref(d), _stream()), "se_conv_gemm")
    return dst


def linear(x2d, W, bias, n_out, act="none", out=None):
    """x2d [M,K] @ W [K, ldw] (+bias, act) -> [M, n_out] through the same implicit-GEMM kernel."""
    m, k = x2d.shape
    assert x2d.is_contiguous()
    if out is None:
        out = torch.empty(m, n_out, device=x2d.device, dtype=torch.float32)
    return conv_gemm(x2d, None, m, 1, 1, 1, [(0, 0)], 1, W, bias, n_out, act, out, 1)


def conv_in1(src, W, bias, cout, act, fout):
    _need_cuda(src, W, bias)
    device_check()
    b, t, fin = src.shape
    assert src.is_contiguous()
    dst = torch.empty(b, t, fout, cout, device=src.device, dtype=torch.float32)
    with _Timed("conv_in1"):
        check(_lib.load().se_conv_in1(_ptr(src), b, t, fin, _ptr(W), _ptr(bias), cout, ACT[act], _ptr(dst), fout,
                                      _stream()), "se_conv_in1")
    return dst


def deconv_out1(src0, src1, W, bias: float, act):
    _need_cuda(src0, src1, W)
    device_check()
    b, t, fin, c0 = src0.shape
    c1 = src1.shape[-1] if src1 is not None else 0
    dst = torch.empty(b, t, 2 * fin + 1, device=src0.device, dtype=torch.float32)
    with _Timed("deconv_out1"):
        check(_lib.load().se_deconv_out1(_ptr(src0), _ptr(src1), c0, c1, b, t, fin, _ptr(W), float(bias), ACT[act],
                                         _ptr(dst), _stream()), "se_deconv_out1")
    return dst


_LSTM_MAX_B = 64


def lstm_seq(xproj, whh, hidden, out=None):
    """xproj [B,T,4H] view (slice-ordered, bias included; rows may be strided), whh packed
    [H/8, H, 32] -> hseq [B,T,H] (out may be a strided view)."""
    _need_cuda(xproj, whh)
    device_check()
    b, t, g4 = xproj.shape
    assert g4 == 4 * hidden and xproj.stride(2) == 1 and xproj.stride(0) == t * xproj.stride(1)
    if out is None:
        out = torch.empty(b, t, hidden, device=xproj.device, dtype=torch.float32)
    lib = _lib.load()
    work = torch.empty(lib.se_lstm_seq_work_bytes(min(b, _LSTM_MAX_B), hidden) // 4, device=xproj.device,
                       dtype=torch.float32)
    sync = torch.zeros(16, device=xproj.device, dtype=torch.int32)
    for b0 in range(0, b, _LSTM_MAX_B):
        nb = min(_LSTM_MAX_B, b - b0)
        xs, os_ = xproj[b0:b0 + nb], out[b0:b0 + nb]
        with _Timed("lstm_seq"):
            check(lib.se_lstm_seq(_ptr(xs), xproj.stride(1), _ptr(whh), nb, t, hidden, _ptr(os_), os_.stride(0),
                                  os_.stride(1), _ptr(work), _ptr(sync), _stream()), "se_lstm_seq")
    return out


def split_tf32(x):
    """x (contiguous fp32) -> (hi, lo) TF32 pair with x ~= hi + lo to 21+ mantissa bits."""
    _need_cuda(x)
    device_check()
    assert x.is_contiguous() and x.numel() % 4 == 0
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with _Timed("split_tf32"):
        check(_lib.load().se_split_tf32(_ptr(x), _ptr(hi), _ptr(lo), x.numel(), _stream()), "se_split_tf32")
    return hi, lo


def pad_split_tf32(x, kpad):
    """x [rows, K] (contiguous fp32) -> (hi, lo) [rows, kpad] TF32 pair, columns K.. zero (see se_pad_split_tf32)."""
    _need_cuda(x)
    device_check()
    rows, k = x.shape
    assert x.is_contiguous() and kpad >= k and kpad % 4 == 0
    hi = torch.empty(rows, kpad, device=x.device, dtype=torch.float32)
    lo = torch.empty_like(hi)
    with _Timed("split_tf32"):
        check(_lib.load().se_pad_split_tf32(_ptr(x), rows, k, kpad, _ptr(hi), _ptr(lo), _stream()), "se_pad_split_tf32")
    return hi, lo


def gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, n_out, act="none", out=None):
    """(a_hi+a_lo) [M,K] @ (b_hi+b_lo) [N,K]^T (+bias, act) -> [M, n_out] on tcgen05 (3xTF32)."""
    _need_cuda(a_hi, a_lo, b_hi, b_lo, bias)
    device_check()
    m, k = a_hi.shape
    n = b_hi.shape[0]
    assert n == n_out and b_hi.shape[1] == k and a_hi.stride(1) == 1 and b_hi.stride(1) == 1
    assert a_lo.stride() == a_hi.stride() and b_lo.stride() == b_hi.stride()
    if out is None:
        out = torch.empty(m, n_out, device=a_hi.device, dtype=torch.float32)
    with _Timed(f"gemm_tf32x3[K={k},N={n}]"):
        check(_lib.load().se_gemm_tf32x3(_ptr(a_hi), _ptr(a_lo), a_hi.stride(0), _ptr(b_hi), _ptr(b_lo),
                                         b_hi.stride(0), m, n, k, _ptr(bias), ACT[act], _ptr(out), out.stride(0),
                                         _stream()), "se_gemm_tf32x3")
    return out


def lstm_cell_tf32x3(x_hi, x_lo, h_hi, h_lo, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out=None):
    """One fused LSTM step for M independent sequences (see se_lstm_cell_tf32x3 in the header)."""
    _need_cuda(x_hi, x_lo, h_hi, h_lo, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out)
    device_check()
    m, kx = x_hi.shape
    hdim = h_hi.shape[1]
    assert w_hi.shape == (4 * hdim, kx + hdim) and x_hi.stride(1) == 1 and h_hi.stride(1) == 1
    assert c_state.is_contiguous() and h_hi_out.is_contiguous() and h_lo_out.is_contiguous()
    with _Timed("lstm_cell_tf32x3"):
        check(_lib.load().se_lstm_cell_tf32x3(_ptr(x_hi), _ptr(x_lo), x_hi.stride(0), kx, _ptr(h_hi), _ptr(h_lo),
                                              h_hi.stride(0), hdim, _ptr(w_hi), _ptr(w_lo), w_hi.stride(0),
                                              _ptr(bias), m, _ptr(c_state), _ptr(h_hi_out), _ptr(h_lo_out),
                                              _ptr(h_out), _stream()), "se_lstm_cell_tf32x3")


def lstm_cell_tf32x3_ex(x_pair, h_pair, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out=None):
    """lstm_cell_tf32x3 with strided h outputs (row stride = .stride(0)) and ``h_pair=None`` for the first step
    (zero initial state; see se_lstm_cell_tf32x3_ex)."""
    x_hi, x_lo = x_pair
    h_hi, h_lo = h_pair if h_pair is not None else (None, None)
    _need_cuda(x_hi, x_lo, h_hi, h_lo, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out)
    device_check()
    m, kx = x_hi.shape
    hdim = h_hi_out.shape[1]
    assert w_hi.shape == (4 * hdim, kx + hdim) and x_hi.stride(1) == 1 and x_lo.stride() == x_hi.stride()
    assert c_state.is_contiguous() and c_state.shape == (m, hdim)
    ldo = h_hi_out.stride(0)
    assert h_hi_out.stride(1) == 1 and h_lo_out.stride() == h_hi_out.stride()
    assert h_out is None or h_out.stride() == h_hi_out.stride()
    assert h_hi is None or (h_hi.stride(1) == 1 and h_lo.stride() == h_hi.stride())
    with _Timed("lstm_cell_tf32x3"):
        check(_lib.load().se_lstm_cell_tf32x3_ex(_ptr(x_hi), _ptr(x_lo), x_hi.stride(0), kx, _ptr(h_hi), _ptr(h_lo),
                                                 h_hi.stride(0) if h_hi is not None else hdim, hdim, _ptr(w_hi),
                                                 _ptr(w_lo), w_hi.stride(0), _ptr(bias), m, _ptr(c_state),
                                                 _ptr(h_hi_out), _ptr(h_lo_out), _ptr(h_out), ldo,
                                                 1 if h_hi is None else 0, _stream()), "se_lstm_cell_tf32x3_ex")


# ---- fp16-pair tensor-core path (kind::f16; see include/se_b200.h "fp16 operand PAIRS") ----------------------------
F16_ACT_SCALE_LOG2 = 4


def _need_cuda_any(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SeB200Error("se_b200 ops need CUDA tensors (no CPU fallback)")


def split_f16(x2d, kpad=None, scale_log2=F16_ACT_SCALE_LOG2):
    """x2d [rows, K] fp32 (row stride free, unit column stride) -> (hi, lo) fp16 [rows, kpad] with
    x * 2^scale_log2 = hi + lo (columns K.. zero)."""
    _need_cuda(x2d)
    device_check()
    rows, k = x2d.shape
    assert x2d.stride(1) == 1
    kpad = kpad or (k + 7) // 8 * 8
    hi = torch.empty(rows, kpad, device=x2d.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    with _Timed("split_f16"):
        check(_lib.load().se_split_f16(_ptr(x2d), rows, k, x2d.stride(0), kpad, scale_log2, _ptr(hi), _ptr(lo),
                                       _stream()), "se_split_f16")
    return hi, lo


def gemm_f16x3(a_pair, b_pair, b_scale_log2, bias, n_out, act="none", out=None, a_scale_log2=F16_ACT_SCALE_LOG2,
               act_param=0.0, alpha=1.0, res=None, want_out=True, pair_out=False, pair16_out=False,
               c16_scale_log2=F16_ACT_SCALE_LOG2):
    """(a_hi+a_lo) [M,K] @ (b_hi+b_lo) [N,Kb]^T (+bias, act) -> [M, n_out] on tcgen05 with fp16 operand pairs.
    Returns out, or (out, (hi, lo) TF32 pair, (hi, lo) fp16 pair) members as requested."""
    a_hi, a_lo = a_pair
    b_hi, b_lo = b_pair
    _need_cuda_any(a_hi, a_lo, b_hi, b_lo)
    _need_cuda(bias, res, out)
    device_check()
    assert a_hi.dtype == torch.float16 and b_hi.dtype == torch.float16
    m, k = a_hi.shape
    n = b_hi.shape[0]
    assert n == n_out and b_hi.shape[1] >= k and a_hi.stride(1) == 1 and b_hi.stride(1) == 1
    assert a_lo.stride() == a_hi.stride() and b_lo.stride() == b_hi.stride()
    dev = a_hi.device
    if out is None and want_out:
        out = torch.empty(m, n_out, device=dev, dtype=torch.float32)
    ldc = out.stride(0) if out is not None else n_out
    c_hi = c_lo = c16_hi = c16_lo = None
    if pair_out:
        c_hi = torch.empty(m, ldc, device=dev, dtype=torch.float32)
        c_lo = torch.empty_like(c_hi)
    if pair16_out:
        c16_hi = torch.empty(m, ldc, device=dev, dtype=torch.float16)
        c16_lo = torch.empty_like(c16_hi)
    with _Timed(f"gemm_f16x3[K={k},N={n}]"):
        check(_lib.load().se_gemm_f16x3(_ptr(a_hi), _ptr(a_lo), a_hi.stride(0), _ptr(b_hi), _ptr(b_lo), b_hi.stride(0),
                                        m, n, k, a_scale_log2 + b_scale_log2, _ptr(bias), ACT[act], float(act_param),
                                        float(alpha), _ptr(res), _ptr(out), _ptr(c_hi), _ptr(c_lo), _ptr(c16_hi),
                                        _ptr(c16_lo), c16_scale_log2, ldc, _stream()), "se_gemm_f16x3")
    if not (pair_out or pair16_out):
        return out
    return out, ((c_hi, c_lo) if pair_out else None), ((c16_hi, c16_lo) if pair16_out else None)


def lstm_cell_f16x3(x_pair, h_pair, cell, c_state, h_hi_out, h_lo_out, h_out=None, a_scale_log2=F16_ACT_SCALE_LOG2):
    """One fused LSTM step for M independent sequences on fp16 operand pairs (se_lstm_cell_f16x3).  ``cell`` is
    packing.pack_lstm_cell_f16's dict; x_pair [M, Kx] and h_pair [M, H] are fp16 pairs scaled by 2^a_scale_log2
    (h_pair=None: zero initial state); the new h leaves in h_hi_out / h_lo_out with the same scale."""
    x_hi, x_lo = x_pair
    h_hi, h_lo = h_pair if h_pair is not None else (None, None)
    _need_cuda_any(x_hi, x_lo, h_hi, h_lo, cell["w_hi"], cell["w_lo"], h_hi_out, h_lo_out)
    _need_cuda(cell["bias"], c_state, h_out)
    device_check()
    m, kx = x_hi.shape
    hdim = cell["hidden"]
    w_hi, w_lo = cell["w_hi"], cell["w_lo"]
    assert x_hi.dtype == torch.float16 and h_hi_out.dtype == torch.float16 and w_hi.dtype == torch.float16
    assert kx <= cell["kx_pad"] and w_hi.shape == (4 * hdim, cell["kx_pad"] + hdim)
    assert x_hi.stride(1) == 1 and x_lo.stride() == x_hi.stride()
    assert c_state.is_contiguous() and c_state.shape == (m, hdim)
    ldo = h_hi_out.stride(0)
    assert h_hi_out.stride(1) == 1 and h_lo_out.stride() == h_hi_out.stride()
    assert h_out is None or h_out.stride() == h_hi_out.stride()
    assert h_hi is None or (h_hi.stride(1) == 1 and h_lo.stride() == h_hi.stride())
    with _Timed("lstm_cell_f16x3"):
        check(_lib.load().se_lstm_cell_f16x3(_ptr(x_hi), _ptr(x_lo), x_hi.stride(0), kx, _ptr(h_hi), _ptr(h_lo),
                                             h_hi.stride(0) if h_hi is not None else hdim, hdim, _ptr(w_hi), _ptr(w_lo),
                                             w_hi.stride(0), a_scale_log2, cell["w_scale_log2"], _ptr(cell["bias"]), m,
                                             _ptr(c_state), _ptr(h_hi_out), _ptr(h_lo_out), _ptr(h_out), ldo,
                                             1 if h_hi is None else 0, _stream()), "se_lstm_cell_f16x3")


def fsn_sb_assemble_f16(mag_tm, fb, nn, inv, scale_log2=F16_ACT_SCALE_LOG2):
    _need_cuda(mag_tm, fb, inv)
    B, Tp, F = mag_tm.shape
    w = 2 * nn + 2
    hi = torch.empty(Tp, B * F, w, device=mag_tm.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    with _Timed("fsn_sb_assemble"):
        check(_lib.load().se_fsn_sb_assemble_f16(_ptr(mag_tm), _ptr(fb), B, Tp, F, nn, _ptr(inv), scale_log2, _ptr(hi),
                                                 _ptr(lo), _stream()), "se_fsn_sb_assemble_f16")
    return hi, lo


def cmul(x, m):
    """x, m [..., 2] interleaved complex -> x * m."""
    _need_cuda(x, m)
    device_check()
    assert x.is_contiguous() and m.is_contiguous() and x.shape == m.shape and x.shape[-1] == 2
    out = torch.empty_like(x)
    with _Timed("cmul"):
        check(_lib.load().se_cmul(_ptr(x), _ptr(m), x.numel() // 2, _ptr(out), _stream()), "se_cmul")
    return out


# ---- FullSubNet glue ----------------------------------------------------------------------------
def fsn_clip_inv_mean(x, strides, B, T, F, denom, wgt=None, extra=None):
    _need_cuda(x, wgt, extra)
    device_check()
    sb, st, sf = strides
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    esb = extra.stride(0) if extra is not None else 0
    ne = extra[0].numel() if extra is not None else 0
    with _Timed("fsn_clip_inv_mean"):
        check(_lib.load().se_fsn_clip_inv_mean(_ptr(x), sb, st, sf, B, T, F, _ptr(wgt), _ptr(extra), esb, ne,
                                               float(denom), _ptr(out), _stream()), "se_fsn_clip_inv_mean")
    return out


def fsn_fb_input(x, strides, B, T, Tp, F, inv):
    _need_cuda(x, inv)
    sb, st, sf = strides
    mag_tm = torch.empty(B, Tp, F, device=x.device, dtype=torch.float32)
    xn = torch.empty_like(mag_tm)
    with _Timed("fsn_fb_input"):
        check(_lib.load().se_fsn_fb_input(_ptr(x), sb, st, sf, B, T, Tp, F, _ptr(inv), _ptr(mag_tm), _ptr(xn),
                                          _stream()), "se_fsn_fb_input")
    return mag_tm, xn


def fsn_sb_assemble(mag_tm, fb, nn, inv):
    _need_cuda(mag_tm, fb, inv)
    B, Tp, F = mag_tm.shape
    w = 2 * nn + 2
    hi = torch.empty(Tp, B * F, w, device=mag_tm.device, dtype=torch.float32)
    lo = torch.empty_like(hi)
    with _Timed("fsn_sb_assemble"):
        check(_lib.load().se_fsn_sb_assemble(_ptr(mag_tm), _ptr(fb), B, Tp, F, nn, _ptr(inv), _ptr(hi), _ptr(lo),
                                             _stream()), "se_fsn_sb_assemble")
    return hi, lo


def fsn_sb_fc(h, W, bias, out):
    _need_cuda(h, W, bias, out)
    m, hd = h.shape
    with _Timed("fsn_sb_fc"):
        check(_lib.load().se_fsn_sb_fc(_ptr(h), m, hd, _ptr(W), _ptr(bias), _ptr(out), _stream()), "se_fsn_sb_fc")
    return out


DCCRN_MASK_MODES = {"E": 0, "C": 1, "R": 2}      # SE_DCCRN_MASK_* in include/se_b200.h


def dccrn_mask(m, x_re, x_im, e_re, e_im, layout_x="btf", layout_e="btf", mode="E"):
    """DCCRN mask application (DCCRN_cprs.py:201-224): 'E' polar, 'C' complex product, 'R' per-component.
    m [B,T,F-1,2] channels-last; x / e planes [B,T,F] ('btf') or [B,F,T]."""
    _need_cuda(m, x_re, x_im, e_re, e_im)
    device_check()
    b, t, fm1, _ = m.shape
    xs = _plane_strides(x_re, layout_x)
    es = _plane_strides(e_re, layout_e)
    assert _plane_strides(x_im, layout_x) == xs and _plane_strides(e_im, layout_e) == es and m.is_contiguous()
    with _Timed("dccrn_mask"):
        check(_lib.load().se_dccrn_mask_ex(_ptr(m), _ptr(x_re), _ptr(x_im), *xs, b, t, fm1 + 1, DCCRN_MASK_MODES[mode],
                                           _ptr(e_re), _ptr(e_im), *es, _stream()), "se_dccrn_mask_ex")


_RESAMPLE_TABLES = {}
_RESAMPLE_TIME = {}


def resample_tables(ratio, device):
    """(win, delta, num_table) device tables of resampy's 'kaiser_best' filter for ``ratio = sr_new / sr_orig``:
    right wing of rolloff * sinc(rolloff * t) on 2**9 samples per zero crossing over 64 zero crossings, tapered by a
    Kaiser window (beta 14.769656459379492, rolloff 0.9475937167399596 -- the parameters resampy documents for the
    table it ships), scaled by ``ratio`` when decimating, and its forward difference (resampy/core.py)."""
    key = (float(ratio), str(device))
    if key not in _RESAMPLE_TABLES:
        import numpy as np
        zeros, bits, rolloff, beta = 64, 512, 0.9475937167399596, 14.769656459379492
        n = bits * zeros
        half = np.kaiser(2 * n + 1, beta)[n:] * rolloff * np.sinc(rolloff * np.linspace(0.0, zeros, n + 1))
        if ratio < 1:
            half = half * ratio
        delta = np.append(np.diff(half), 0.0)
        _RESAMPLE_TABLES[key] = (torch.from_numpy(half.astype(np.float32)).to(device),
                                 torch.from_numpy(delta.astype(np.float32)).to(device), bits)
    return _RESAMPLE_TABLES[key]


def resample(x, sr_orig, sr_new, out=None):
    """librosa.resample(x, sr_orig, sr_new, fix=True, scale=False) for a batch (LSTM/lstm_decode_vb.py:34).
    x [B, n_in] float32 CUDA -> [B, ceil(n_in * sr_new / sr_orig)]; identity when the rates agree."""
    _need_cuda(x)
    device_check()
    if int(sr_orig) == int(sr_new):
        return x
    assert x.dim() == 2 and x.stride(1) == 1
    ratio = float(sr_new) / float(sr_orig)
    b, n_in = x.shape
    n_valid = int(n_in * ratio)                       # resampy.core.resample: int(shape * sample_ratio)
    n_out = int(math.ceil(n_in * ratio))              # librosa: fix_length(y_hat, ceil(n * ratio))
    if n_valid < 1:
        raise ValueError("input too short to resample")
    win, delta, bits = resample_tables(ratio, x.device)
    tkey = (float(ratio), n_valid, str(x.device))
    if tkey not in _RESAMPLE_TIME:           # resample_f's time register: 0, += 1/ratio (sequential float64 adds)
        import numpy as np
        treg = np.concatenate([[0.0], np.cumsum(np.full(n_valid - 1, 1.0 / ratio, dtype=np.float64))])
        while len(_RESAMPLE_TIME) >= 4:      # small cache: corpora are processed group by group; evicted tables stay
            _RESAMPLE_TIME.pop(next(iter(_RESAMPLE_TIME)))   # referenced by the caching allocator's stream ordering
        _RESAMPLE_TIME[tkey] = torch.from_numpy(treg).to(x.device)
    treg = _RESAMPLE_TIME[tkey]
    if out is None:
        out = torch.empty(b, n_out, device=x.device, dtype=torch.float32)
    with _Timed("resample"):
        check(_lib.load().se_resample(_ptr(x), x.stride(0), b, n_in, _ptr(out), out.stride(0), n_out, n_valid, ratio,
                                      _ptr(treg), _ptr(win), _ptr(delta), win.numel(), bits, _stream()), "se_resample")
    return out


def conv_tf32x3(src0, src1, B, T, Fin, Fout, taps, sf, w_hi, w_lo, bias, Cout, act, dstF, dst_f0=0, dst_fstep=1,
                act_param=0.0, out=None, out_pair=None, glu=None):
    """Tensor-core implicit-GEMM conv.  src0 / src1: (hi, lo) tuples of channels-last [B,T,Fin,C];
    w_hi / w_lo [Cout, ntaps*(C0+C1)].  out: fp32 [B,T,dstF,Cout] or None; out_pair: (hi, lo) or None.
    glu = (scale, shift) (either may be None): gated conv, GEMM columns are (conv1, conv2) pairs and the outputs have
    Cout / 2 channels (see se_conv_tc_desc)."""
    device_check()
    d = ConvTcDesc()
    d.src0_hi, d.src0_lo = src0[0].data_ptr(), src0[1].data_ptr()
    c0 = src0[0].shape[-1]
    c1 = src1[0].shape[-1] if src1 is not None else 0
    d.src1_hi = src1[0].data_ptr() if src1 is not None else 0
    d.src1_lo = src1[1].data_ptr() if src1 is not None else 0
    d.C0, d.C1, d.B, d.T, d.Fin, d.Fout = c0, c1, B, T, Fin, Fout
    d.ntaps = len(taps)
    for i, (dt, df) in enumerate(taps):
        d.dt[i], d.df[i] = dt, df
    d.sf = sf
    assert w_hi.shape == (Cout, len(taps) * (c0 + c1)) and w_hi.is_contiguous() and w_lo.is_contiguous()
    d.w_hi, d.w_lo = w_hi.data_ptr(), w_lo.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else 0
    d.Cout, d.act, d.act_param = Cout, ACT[act], float(act_param)
    d.out = out.data_ptr() if out is not None else 0
    d.out_hi = out_pair[0].data_ptr() if out_pair is not None else 0
    d.out_lo = out_pair[1].data_ptr() if out_pair is not None else 0
    d.dstF, d.dst_f0, d.dst_fstep = dstF, dst_f0, dst_fstep
    d.glu = 0 if glu is None else 1
    d.glu_scale = glu[0].data_ptr() if glu is not None and glu[0] is not None else 0
    d.glu_shift = glu[1].data_ptr() if glu is not None and glu[1] is not None else 0
    with _Timed(f"conv_tf32x3[K={len(taps) * (c0 + c1)},N={Cout}]"):
        check(_lib.load().se_conv_tf32x3(C.byref(d), _stream()), "se_conv_tf32x3")


def conv_f16x3(src0, src1, B, T, Fin, Fout, taps, sf, w_hi, w_lo, w_scale_log2, bias, Cout, act, dstF, dst_f0=0,
               dst_fstep=1, act_param=0.0, out=None, out_pair=None, out_pair16=None, glu=None,
               a_scale_log2=F16_ACT_SCALE_LOG2, out16_scale_log2=F16_ACT_SCALE_LOG2, fout1=None):
    """conv_tf32x3 on fp16 operand pairs (se_conv_f16x3).  src0 / src1: (hi, lo) fp16 tuples of channels-last
    [B,T,Fin,C] scaled by 2^a_scale_log2; w_hi / w_lo fp16 [Cout, ntaps*(pad64(C0)+pad64(C1))] scaled by 2^w_scale_log2
    (packing.pack_conv_f16).  Outputs as requested: fp32, TF32 pair, fp16 pair (scaled by 2^out16_scale_log2).
    fout1 is not None: two output-column parity classes in one launch (se_conv_f16_desc.ncls = 2): Cout counts the
    GEMM columns of both classes, the outputs have Cout / 2 channels."""
    device_check()
    d = ConvF16Desc()
    assert src0[0].dtype == torch.float16 and w_hi.dtype == torch.float16
    d.src0_hi, d.src0_lo = src0[0].data_ptr(), src0[1].data_ptr()
    c0 = src0[0].shape[-1]
    c1 = src1[0].shape[-1] if src1 is not None else 0
    d.src1_hi = src1[0].data_ptr() if src1 is not None else 0
    d.src1_lo = src1[1].data_ptr() if src1 is not None else 0
    d.C0, d.C1, d.B, d.T, d.Fin, d.Fout = c0, c1, B, T, Fin, Fout
    d.ntaps = len(taps)
    for i, (dt, df) in enumerate(taps):
        d.dt[i], d.df[i] = dt, df
    d.sf = sf
    kpad = len(taps) * ((c0 + 63) // 64 * 64 + (c1 + 63) // 64 * 64)
    assert w_hi.shape == (Cout, kpad) and w_hi.is_contiguous() and w_lo.is_contiguous(), (w_hi.shape, Cout, kpad)
    d.w_hi, d.w_lo = w_hi.data_ptr(), w_lo.data_ptr()
    d.scale_log2_a, d.scale_log2_w = a_scale_log2, w_scale_log2
    d.bias = bias.data_ptr() if bias is not None else 0
    d.Cout, d.act, d.act_param = Cout, ACT[act], float(act_param)
    d.out = out.data_ptr() if out is not None else 0
    d.out_hi = out_pair[0].data_ptr() if out_pair is not None else 0
    d.out_lo = out_pair[1].data_ptr() if out_pair is not None else 0
    d.out16_hi = out_pair16[0].data_ptr() if out_pair16 is not None else 0
    d.out16_lo = out_pair16[1].data_ptr() if out_pair16 is not None else 0
    d.out16_scale_log2 = out16_scale_log2
    d.dstF, d.dst_f0, d.dst_fstep = dstF, dst_f0, dst_fstep
    d.glu = 0 if glu is None else 1
    d.glu_scale = glu[0].data_ptr() if glu is not None and glu[0] is not None else 0
    d.glu_shift = glu[1].data_ptr() if glu is not None and glu[1] is not None else 0
    d.ncls, d.fout1 = (2, fout1) if fout1 is not None else (0, 0)
    with _Timed(f"conv_f16x3[K={len(taps) * (c0 + c1)},N={Cout}]"):
        check(_lib.load().se_conv_f16x3(C.byref(d), _stream()), "se_conv_f16x3")


def fill_column(dst, fill, fill_f, act, act_param=0.0):
    """dst [B,T,F,C]: column fill_f <- act(fill[c])."""
    _need_cuda(dst, fill)
    b, t, f, c = dst.shape
    with _Timed("fill_column"):
        check(_lib.load().se_fill_column(_ptr(dst), b * t, f, c, fill_f, _ptr(fill), ACT[act], float(act_param),
                                         _stream()), "se_fill_column")


def lstm_seq_multi(xproj, whh, hidden, ngroups, out):
    """``ngroups`` same-shape LSTMs in one launch.  xproj [B,T,ngroups*4H] (group-major column blocks),
    whh [ngroups, H/8, H, 32] -- or [H/8, H, 32] when all groups share one weight set (DPCRN's inter-LSTM:
    the groups are the F = 4 frequency positions) --, out [B,T,ngroups*H]."""
    _need_cuda(xproj, whh, out)
    device_check()
    b, t, cols = xproj.shape
    assert cols == ngroups * 4 * hidden and xproj.is_contiguous() and out.is_contiguous() and whh.is_contiguous()
    assert out.shape == (b, t, ngroups * hidden)
    shared = whh.dim() == 3
    wstride = 0 if shared else whh[0].numel()
    lib = _lib.load()
    work = torch.empty(ngroups * lib.se_lstm_seq_work_bytes(min(b, _LSTM_MAX_B), hidden) // 4, device=xproj.device,
                       dtype=torch.float32)
    sync = torch.zeros(16, device=xproj.device, dtype=torch.int32)
    for b0 in range(0, b, _LSTM_MAX_B):
        nb = min(_LSTM_MAX_B, b - b0)
        xs, os_ = xproj[b0:b0 + nb], out[b0:b0 + nb]
        with _Timed("lstm_seq_multi"):
            check(lib.se_lstm_seq_multi(_ptr(xs), cols, 4 * hidden, _ptr(whh), wstride, ngroups, nb, t, hidden,
                                        _ptr(os_), os_.stride(0), os_.stride(1), hidden, _ptr(work), _ptr(sync),
                                        _stream()), "se_lstm_seq_multi")
    return out


def gemm_tf32x3_ex(a_pair, b_hi, b_lo, bias, n_out, act="none", act_param=0.0, alpha=1.0, res=None, want_f32=True,
                   want_pair=False):
    """Extended tensor-core GEMM: returns (C or None, (c_hi, c_lo) or None); C = alpha*act(A B^T + bias) + res."""
    a_hi, a_lo = a_pair
    _need_cuda(a_hi, a_lo, b_hi, b_lo, bias, res)
    device_check()
    m, k = a_hi.shape
    assert b_hi.shape == (n_out, k) and a_hi.stride(1) == 1 and a_lo.stride() == a_hi.stride()
    mk = lambda: torch.empty(m, n_out, device=a_hi.device, dtype=torch.float32)   # noqa: E731
    out = mk() if want_f32 else None
    pair = (mk(), mk()) if want_pair else None
    if res is not None:
        assert res.shape == (m, n_out) and res.is_contiguous()
    with _Timed(f"gemm_tf32x3[K={k},N={n_out}]"):
        check(_lib.load().se_gemm_tf32x3_ex(_ptr(a_hi), _ptr(a_lo), a_hi.stride(0), _ptr(b_hi), _ptr(b_lo),
                                            b_hi.stride(0), m, n_out, k, _ptr(bias), ACT[act], float(act_param),
                                            float(alpha), _ptr(res), _ptr(out), _ptr(pair[0] if pair else None),
                                            _ptr(pair[1] if pair else None), n_out, _stream()), "se_gemm_tf32x3_ex")
    return out, pair


# ---- Uformer glue -----------------------------------------------------------------------------------
def uf_prep(x):
    """x [B,T,F,2] -> mag [B,T,F], phase [B,T,F], cplx_in [B,T,F-1,2], mag_in [B,T,F-1,1]."""
    _need_cuda(x)
    device_check()
    b, t, f, _ = x.shape
    assert x.is_contiguous()
    mk = lambda *s: torch.empty(*s, device=x.device, dtype=torch.float32)   # noqa: E731
    mag, phase, cin, min_ = mk(b, t, f), mk(b, t, f), mk(b, t, f - 1, 2), mk(b, t, f - 1, 1)
    with _Timed("uf_prep"):
        check(_lib.load().se_uf_prep(_ptr(x), b, t, f, _ptr(mag), _ptr(phase), _ptr(cin), _ptr(min_), _stream()),
              "se_uf_prep")
    return mag, phase, cin, min_


def uf_fusion(c, m):
    """c [..., 2C], m [..., C] -> (c', m') (fusion.py:13-19)."""
    _need_cuda(c, m)
    device_check()
    ch = m.shape[-1]
    rows = m.numel() // ch
    assert c.shape[-1] == 2 * ch and c.is_contiguous() and m.is_contiguous()
    co, mo = torch.empty_like(c), torch.empty_like(m)
    with _Timed("uf_fusion"):
        check(_lib.load().se_uf_fusion(_ptr(c), _ptr(m), rows, ch, _ptr(co), _ptr(mo), _stream()), "se_uf_fusion")
    return co, mo


def uf_fusion_ex(c, m, c_f32=True, c_pair=False, m_f32=True, m_pair=False):
    """uf_fusion with every result as fp32 and / or TF32 (hi, lo) pair: returns (c_f32 | None, c_pair | None,
    m_f32 | None, m_pair | None)."""
    _need_cuda(c, m)
    device_check()
    ch = m.shape[-1]
    rows = m.numel() // ch
    assert c.shape[-1] == 2 * ch and c.is_contiguous() and m.is_contiguous()
    assert (c_f32 or c_pair) and (m_f32 or m_pair)
    co = torch.empty_like(c) if c_f32 else None
    mo = torch.empty_like(m) if m_f32 else None
    cp = (torch.empty_like(c), torch.empty_like(c)) if c_pair else None
    mp = (torch.empty_like(m), torch.empty_like(m)) if m_pair else None
    with _Timed("uf_fusion"):
        check(_lib.load().se_uf_fusion_ex(_ptr(c), _ptr(m), rows, ch, _ptr(co), _ptr(cp[0] if cp else None),
                                          _ptr(cp[1] if cp else None), _ptr(mo), _ptr(mp[0] if mp else None),
                                          _ptr(mp[1] if mp else None), _stream()), "se_uf_fusion_ex")
    return co, cp, mo, mp


_LN_POST = {"none": 0, "prelu": 1, "swish": 2}


def group_layernorm(x, groups, gamma, beta, gate=None, post="none", slope=0.0, res=None, want_f32=True,
                    want_pair=False, eps=1e-5, out_index=None):
    """LayerNorm over each of ``groups`` channel groups of the last axis (see se_group_layernorm).
    out_index (int32 [C]): channel ch is stored at position out_index[ch] of its group."""
    _need_cuda(x, gamma, beta, gate, res)
    if out_index is not None and not out_index.is_cuda:
        raise _lib.SeB200Error("se_b200 ops need CUDA tensors (no CPU fallback)")
    assert out_index is None or (out_index.dtype == torch.int32 and out_index.numel() == x.shape[-1] // groups)
    device_check()
    ctot = x.shape[-1]
    c = ctot // groups
    rows = x.numel() // ctot
    assert x.is_contiguous() and gamma.numel() == c
    out = torch.empty_like(x) if want_f32 else None
    pair = (torch.empty_like(x), torch.empty_like(x)) if want_pair else None
    with _Timed(f"group_layernorm[C={c}]"):
        check(_lib.load().se_group_layernorm(_ptr(x), _ptr(gate), rows, groups, c, _ptr(gamma), _ptr(beta), float(eps),
                                             _LN_POST[post], float(slope), _ptr(res), _ptr(out_index), _ptr(out),
                                             _ptr(pair[0] if pair else None), _ptr(pair[1] if pair else None),
                                             _stream()), "se_group_layernorm")
    return out, pair


def attention(qkv, nheads, head_out, head_sign, nout, L, lstride, n_outer, outer_stride, n_inner, inner_stride,
              scale=0.25):
    """qkv [R, nheads*48] -> out [R, nout*16]."""
    _need_cuda(qkv)
    device_check()
    r, ld = qkv.shape
    assert qkv.is_contiguous() and ld == nheads * 48
    out = torch.empty(r, nout * 16, device=qkv.device, dtype=torch.float32)
    ho = (C.c_int * 8)(*(list(head_out) + [0] * (8 - len(head_out))))
    hs = (C.c_float * 8)(*(list(head_sign) + [0.0] * (8 - len(head_sign))))
    with _Timed(f"attention[L={L},heads={nheads}]"):
        check(_lib.load().se_attention(_ptr(qkv), ld, nheads, ho, hs, nout, L, lstride, n_outer, outer_stride, n_inner,
                                       inner_stride, float(scale), _ptr(out), nout * 16, _stream()), "se_attention")
    return out


def uf_mask(cmask, mdec, mag, phase):
    """cmask [B,T,F-1,2], mdec [B,T,F-1,1], mag/phase [B,T,F] -> est [B,T,F,2]."""
    _need_cuda(cmask, mdec, mag, phase)
    device_check()
    b, t, f = mag.shape
    est = torch.empty(b, t, f, 2, device=mag.device, dtype=torch.float32)
    with _Timed("uf_mask"):
        check(_lib.load().se_uf_mask(_ptr(cmask), _ptr(mdec), _ptr(mag), _ptr(phase), b, t, f, _ptr(est), _stream()),
              "se_uf_mask")
    return est


def set_lstm_engine(engine: int):
    """0 = fp32 FMA recurrence kernel, 1 = mma.sync 3xTF32 kernel, 2 = tcgen05 3xTF32 cluster kernel (H = 1024), 3 = 2 plus
    the sequence-parallel kernel at H = 128, 4 (default) = 3 with the fp16-pair / tagged-state tcgen05 kernel at H = 1024,
    5 = 4 with two independent half-batch chains per CTA."""
    check(_lib.load().se_set_lstm_engine(int(engine)), "se_set_lstm_engine")


def set_gemm_engine(engine: int):
    """0 = one CTA per 128x128 tile, 1 = CTA pairs (cta_group::2, 256x256 tiles) where M, N >= 256, 2 / 3 / 4 = multicast
    clusters of 2x2 / 1x2 / 2x1 CTAs (see se_set_gemm_engine in the header)."""
    check(_lib.load().se_set_gemm_engine(int(engine)), "se_set_gemm_engine")


def glu_affine_act(x, scale, shift, act="elu", act_param=0.0, want_f32=True, want_pair=False):
    """x [..., 2C] = [a | b] -> act((a * sigmoid(b)) * scale + shift) [..., C]."""
    _need_cuda(x, scale, shift)
    device_check()
    c = x.shape[-1] // 2
    rows = x.numel() // (2 * c)
    assert x.is_contiguous()
    mk = lambda: torch.empty(*x.shape[:-1], c, device=x.device, dtype=torch.float32)   # noqa: E731
    out = mk() if want_f32 else None
    pair = (mk(), mk()) if want_pair else None
    with _Timed("glu_affine_act"):
        check(_lib.load().se_glu_affine_act(_ptr(x), rows, c, _ptr(scale), _ptr(shift), ACT[act], float(act_param),
                                            _ptr(out), _ptr(pair[0] if pair else None),
                                            _ptr(pair[1] if pair else None), _stream()), "se_glu_affine_act")
    return out, pair


def unary(x, act, act_param=0.0, want_f32=True, want_pair=False):
    _need_cuda(x)
    device_check()
    assert x.is_contiguous()
    out = torch.empty_like(x) if want_f32 else None
    pair = (torch.empty_like(x), torch.empty_like(x)) if want_pair else None
    with _Timed("unary"):
        check(_lib.load().se_unary(_ptr(x), x.numel(), ACT[act], float(act_param), _ptr(out),
                                   _ptr(pair[0] if pair else None), _ptr(pair[1] if pair else None), _stream()),
              "se_unary")
    return out, pair


# ---- utterance-level norms of the TCM family (csrc/norm.cu) ------------------------------------------
_norm_ws = {}    # device -> zero-initialised workspace (se_chan_stats leaves it zeroed)


def _stats_ws(device, nbytes):
    ws = _norm_ws.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, device=device, dtype=torch.uint8)
        _norm_ws[device] = ws
    return ws


def chan_stats(x, B, rows, C, pre="none", pre_slope=None, eps=1e-5):
    """InstanceNorm statistics of pre(x) per (clip, channel).  x [B, rows, Cin] -> (mean [B,C], rstd [B,C])."""
    _need_cuda(x, pre_slope)
    device_check()
    assert x.is_contiguous() and x.numel() % (B * rows) == 0
    cin = x.numel() // (B * rows)
    lib = _lib.load()
    ws = _stats_ws(x.device, int(lib.se_chan_stats_ws_bytes(B, rows, C)))
    mean = torch.empty(B, C, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    with _Timed(f"chan_stats[C={C}]"):
        check(lib.se_chan_stats(_ptr(x), B, rows, cin, C, _lib.NORM_PRE[pre], _ptr(pre_slope), float(eps), _ptr(mean),
                                _ptr(rstd), _ptr(ws), _stream()), "se_chan_stats")
    return mean, rstd


def cum_stats(x, B, T, F, C, pre="none", pre_slope=None, eps=1e-5, groups=1):   # noqa: N803 (C shadows the ctypes alias)
    """Cumulative-LayerNorm statistics of pre(x): x [B, T, F, Cin] -> (mean [B,T,G], rstd [B,T,G])."""
    _need_cuda(x, pre_slope)
    device_check()
    assert x.is_contiguous() and x.numel() % (B * T * F) == 0
    cin = x.numel() // (B * T * F)
    ws = torch.empty(2 * B * T * groups, device=x.device, dtype=torch.float64)
    mean = torch.empty(B, T, groups, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    with _Timed(f"cum_stats[C={C}]"):
        check(_lib.load().se_cum_stats(_ptr(x), B, T, F, cin, C, groups, _lib.NORM_PRE[pre], _ptr(pre_slope), float(eps),
                                       _ptr(mean), _ptr(rstd), _ptr(ws), _stream()), "se_cum_stats")
    return mean, rstd


def chan_norm(x, B, rows, C, mean, rstd, gamma, beta, pre="none", pre_slope=None, cumulative=False, rows_per_t=1,
              post="none", post_slope=None, fir_w=None, fir_groups=1, want_f32=True, want_pair=False, stat_groups=1):
    """post((pre(x) - mean) * rstd * gamma + beta); x [B, rows, Cin] -> [B, rows, C] fp32 and/or TF32 pair."""
    _need_cuda(x, pre_slope, mean, rstd, gamma, beta, post_slope, fir_w)
    device_check()
    assert x.is_contiguous() and x.numel() % (B * rows) == 0
    cin = x.numel() // (B * rows)
    mk = lambda: torch.empty(B, rows, C, device=x.device, dtype=torch.float32)   # noqa: E731
    out = mk() if want_f32 else None
    pair = (mk(), mk()) if want_pair else None
    fir_k = 0
    if fir_w is not None:
        assert fir_w.is_contiguous() and fir_w.dim() == 2 and fir_w.shape[0] == fir_groups
        fir_k = fir_w.shape[1]
    with _Timed(f"chan_norm[C={C},{post}]"):
        check(_lib.load().se_chan_norm(_ptr(x), B, rows, cin, C, _lib.NORM_PRE[pre], _ptr(pre_slope), _ptr(mean),
                                       _ptr(rstd), 1 if cumulative else 0, rows_per_t, stat_groups, _ptr(gamma), _ptr(beta),
                                       _lib.NORM_POST[post], _ptr(post_slope), _ptr(fir_w), fir_k, fir_groups,
                                       _ptr(out), _ptr(pair[0] if pair else None), _ptr(pair[1] if pair else None),
                                       _stream()), "se_chan_norm")
    return out, pair


def add(a, b, want_f32=True, want_pair=False):
    _need_cuda(a, b)
    device_check()
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    out = torch.empty_like(a) if want_f32 else None
    pair = (torch.empty_like(a), torch.empty_like(a)) if want_pair else None
    with _Timed("add"):
        check(_lib.load().se_add(_ptr(a), _ptr(b), a.numel(), _ptr(out), _ptr(pair[0] if pair else None),
                                 _ptr(pair[1] if pair else None), _stream()), "se_add")
    return out, pair


def axpby(a, b, ca, cb, want_f32=True, want_pair=False):
    """ca * a + cb * b."""
    _need_cuda(a, b)
    device_check()
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    out = torch.empty_like(a) if want_f32 else None
    pair = (torch.empty_like(a), torch.empty_like(a)) if want_pair else None
    with _Timed("axpby"):
        check(_lib.load().se_axpby(_ptr(a), _ptr(b), float(ca), float(cb), a.numel(), _ptr(out),
                                   _ptr(pair[0] if pair else None), _ptr(pair[1] if pair else None), _stream()),
              "se_axpby")
    return out, pair


def taylor_zero(x_ri, gain, ld, want_pair=True):
    """x_ri [B,T,F,2], gain [B,T,F] -> zeroth-order term as RI rows [B*T, ld] (fp32, TF32 pair)."""
    _need_cuda(x_ri, gain)
    device_check()
    assert x_ri.is_contiguous() and gain.is_contiguous() and x_ri.shape[:-1] == gain.shape
    f = gain.shape[-1]
    rows = gain.numel() // f
    out = torch.empty(rows, ld, device=gain.device, dtype=torch.float32)
    pair = (torch.empty_like(out), torch.empty_like(out)) if want_pair else None
    with _Timed("taylor_zero"):
        check(_lib.load().se_taylor_zero(_ptr(x_ri), _ptr(gain), rows, f, ld, _ptr(out), _ptr(pair[0] if pair else None),
                                         _ptr(pair[1] if pair else None), _stream()), "se_taylor_zero")
    return out, pair


def gaf_update(x_re, x_im, xs_r, xs_f, gain, resi, rows, f, ld, im_off, want_f32=True, want_pair=True):
    """G2Net stage update on RI rows (see se_gaf_update).  x_re / x_im: tensors whose data_ptr is element (0, 0) of the
    real / imaginary plane of ``pre``; gain [rows, F] and resi [rows, ld] or both None (relayout only)."""
    _need_cuda(x_re, x_im, gain, resi)
    device_check()
    out = torch.empty(rows, ld, device=x_re.device, dtype=torch.float32) if want_f32 else None
    pair = (torch.empty(rows, ld, device=x_re.device, dtype=torch.float32),
            torch.empty(rows, ld, device=x_re.device, dtype=torch.float32)) if want_pair else None
    with _Timed("gaf_update"):
        check(_lib.load().se_gaf_update(_ptr(x_re), _ptr(x_im), xs_r, xs_f, _ptr(gain), _ptr(resi), rows, f, ld, im_off,
                                        _ptr(out), _ptr(pair[0] if pair else None), _ptr(pair[1] if pair else None),
                                        _stream()), "se_gaf_update")
    return out, pair


def cts_glue1(x_ri, est_mag):
    """x_ri [B,T,F,2] noisy (compressed) RI, est_mag [B,T,F] -> s2_in [B,T,F,4] (two_stage_com_decode_vb.py:79-82)."""
    _need_cuda(x_ri, est_mag)
    device_check()
    assert x_ri.is_contiguous() and est_mag.is_contiguous() and x_ri.shape[:-1] == est_mag.shape
    out = torch.empty(*est_mag.shape, 4, device=x_ri.device, dtype=torch.float32)
    with _Timed("cts_glue1"):
        check(_lib.load().se_cts_glue1(_ptr(x_ri), _ptr(est_mag), est_mag.numel(), _ptr(out), _stream()), "se_cts_glue1")
    return out


def cts_glue2(out_r, out_i, s2_in):
    """stage-2 output + stage-1 RI (two_stage_com_decode_vb.py:84) -> est [B,T,F,2]."""
    _need_cuda(out_r, out_i, s2_in)
    device_check()
    assert out_r.is_contiguous() and out_i.is_contiguous() and s2_in.is_contiguous()
    est = torch.empty(*out_r.shape, 2, device=out_r.device, dtype=torch.float32)
    with _Timed("cts_glue2"):
        check(_lib.load().se_cts_glue2(_ptr(out_r), _ptr(out_i), _ptr(s2_in), out_r.numel(), _ptr(est), _stream()),
              "se_cts_glue2")
    return est
