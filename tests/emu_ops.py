"""Pure-torch mirror of the C-ABI *semantics* (include/se_b200.h), used ONLY by the CPU tests to
check the host-side logic (weight packing, tap tables, layouts, orchestration) without a GPU.
It is not a fallback: product code never imports it; tests monkeypatch ``ops`` with it and call
``model._forward_impl`` directly."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _act(x, act, param=0.0):
    if act == "prelu":
        return torch.where(x >= 0, x, param * x)
    return {"none": lambda v: v, "elu": F.elu, "softplus": F.softplus, "relu": F.relu,
            "sigmoid": torch.sigmoid, "tanh": torch.tanh}[act](x)


def conv_gemm(src0, src1, B, T, Fin, Fout, taps, sf, W, bias, Cout, act, dst, dstF, dst_f0=0, dst_fstep=1,
              fill_f=-1, fill=None, act_param=0.0):
    s0 = src0.reshape(B, T, Fin, -1)
    x = s0 if src1 is None else torch.cat([s0, src1.reshape(B, T, Fin, -1)], dim=-1)
    ct = x.shape[-1]
    acc = torch.zeros(B, T, Fout, W.shape[1], dtype=x.dtype)
    fo = torch.arange(Fout)
    for i, (dt, df) in enumerate(taps):
        fi = fo * sf + df
        okf = (fi >= 0) & (fi < Fin)
        g = torch.zeros(B, T, Fout, ct, dtype=x.dtype)
        tsrc = torch.arange(T) + dt
        okt = (tsrc >= 0) & (tsrc < T)
        sel = x[:, tsrc.clamp(0, T - 1)][:, :, fi.clamp(0, Fin - 1)]
        sel = sel * okt[None, :, None, None] * okf[None, None, :, None]
        acc += sel @ W[i * ct:(i + 1) * ct]
    out = acc[..., :Cout]
    if bias is not None:
        out = out + bias
    out = _act(out, act, act_param)
    d = dst.view(B, T, dstF, Cout)
    d[:, :, dst_f0:dst_f0 + (Fout - 1) * dst_fstep + 1:dst_fstep] = out
    if fill_f >= 0:
        d[:, :, fill_f] = _act(fill, act, act_param)
    return dst


def linear(x2d, W, bias, n_out, act="none", out=None):
    y = x2d @ W[:, :n_out]
    if bias is not None:
        y = y + bias
    return _act(y, act)


def conv_in1(src, W, bias, cout, act, fout):
    b, t, fin = src.shape
    dst = torch.empty(b, t, fout, cout, dtype=src.dtype)
    taps = [(kt - 1, kf) for kt in range(2) for kf in range(3)]
    return conv_gemm(src.unsqueeze(-1), None, b, t, fin, fout, taps, 2, W, bias, cout, act, dst, fout)


def deconv_out1(src0, src1, W, bias, act):
    b, t, fin, c0 = src0.shape
    x = src0 if src1 is None else torch.cat([src0, src1], dim=-1)
    ct = x.shape[-1]
    out = torch.full((b, t, 2 * fin + 1), float(bias), dtype=src0.dtype)
    for kt in range(2):
        xs = torch.zeros_like(x)
        xs[:, kt:] = x[:, :t - kt] if kt else x
        for kf in range(3):
            contrib = xs @ W[kt * 3 + kf]            # [b,t,fin]
            out[:, :, kf:kf + 2 * fin:2] += contrib
    return _act(out, act)


def lstm_seq(xproj, whh, hidden, out=None):
    b, t, _ = xproj.shape
    s = hidden // 8
    h = torch.zeros(b, hidden, dtype=xproj.dtype)
    c = torch.zeros(b, hidden, dtype=xproj.dtype)
    outs = []
    # whh [S, H, 32]: gates for slice s = h @ whh[s]  -> [b, 32] = (gate, j)
    for step in range(t):
        g = xproj[:, step].view(b, s, 4, 8) + torch.einsum("bk,skn->bsn", h, whh).view(b, s, 4, 8)
        i, f, gg, o = g[:, :, 0], g[:, :, 1], g[:, :, 2], g[:, :, 3]
        c = (torch.sigmoid(f) * c.view(b, s, 8) + torch.sigmoid(i) * torch.tanh(gg)).reshape(b, hidden)
        h = (torch.sigmoid(o).reshape(b, hidden) * torch.tanh(c))
        outs.append(h)
    res = torch.stack(outs, dim=1)
    if out is not None:
        out.copy_(res)
        return out
    return res


def lstm_seq_multi(xproj, whh, hidden, ngroups, out):
    for g in range(ngroups):
        w = whh if whh.dim() == 3 else whh[g]
        out[:, :, g * hidden:(g + 1) * hidden] = lstm_seq(xproj[:, :, g * 4 * hidden:(g + 1) * 4 * hidden], w, hidden)
    return out


def dccrn_mask(m, x_re, x_im, e_re, e_im, layout_x="btf", layout_e="btf", mode="E"):
    assert layout_x == "btf" and layout_e == "btf"
    M = torch.view_as_complex(m.contiguous())
    X = torch.complex(x_re, x_im)
    E = torch.zeros_like(X)
    if mode == "E":
        mm = M.abs()
        g = torch.where(mm > 0, torch.tanh(mm) / mm.clamp_min(1e-30), torch.zeros_like(mm))
        E[:, :, 1:] = X[:, :, 1:] * M * g
    elif mode == "C":
        E[:, :, 1:] = X[:, :, 1:] * M
    else:
        E[:, :, 1:] = torch.complex(X.real[:, :, 1:] * M.real, X.imag[:, :, 1:] * M.imag)
    e_re.copy_(E.real)
    e_im.copy_(E.imag)


def conv_tf32x3(src0, src1, B, T, Fin, Fout, taps, sf, w_hi, w_lo, bias, Cout, act, dstF, dst_f0=0, dst_fstep=1,
                act_param=0.0, out=None, out_pair=None, glu=None):
    from se_b200 import packing
    x0 = src0[0] + src0[1]
    x1 = (src1[0] + src1[1]) if src1 is not None else None
    w = (w_hi + w_lo).t().contiguous()
    if glu is not None:      # gated: columns (2j, 2j+1) = (conv1, conv2) of channel j, outputs have Cout / 2 channels
        full = torch.zeros(B, T, dstF, Cout, dtype=x0.dtype)
        conv_gemm(x0, x1, B, T, Fin, Fout, taps, sf, w, bias, Cout, "none", full, dstF, dst_f0, dst_fstep, -1, None, 0.0)
        fs = torch.arange(Fout) * dst_fstep + dst_f0
        v = full[:, :, fs, 0::2] * torch.sigmoid(full[:, :, fs, 1::2])
        if glu[0] is not None:
            v = v * glu[0]
        if glu[1] is not None:
            v = v + glu[1]
        v = _act(v, act, act_param)
        if out is not None:
            out[:, :, fs] = v
        if out_pair is not None:
            hi, lo = packing.split_tf32(v.float().contiguous())
            out_pair[0][:, :, fs] = hi
            out_pair[1][:, :, fs] = lo
        return
    tmp = out if out is not None else torch.zeros(B, T, dstF, Cout, dtype=x0.dtype)
    if out is None and out_pair is not None:
        tmp.copy_(out_pair[0] + out_pair[1])
    conv_gemm(x0, x1, B, T, Fin, Fout, taps, sf, w, bias, Cout, act, tmp, dstF, dst_f0, dst_fstep, -1, None, act_param)
    if out_pair is not None:
        hi, lo = packing.split_tf32(tmp)
        out_pair[0].copy_(hi)
        out_pair[1].copy_(lo)


def fill_column(dst, fill, fill_f, act, act_param=0.0):
    dst[:, :, fill_f] = _act(fill, act, act_param)


def split_tf32(x):
    from se_b200 import packing
    return packing.split_tf32(x)


def pad_split_tf32(x, kpad):
    from se_b200 import packing
    xp = torch.cat([x, x.new_zeros(x.shape[0], kpad - x.shape[1])], dim=1)
    return packing.split_tf32(xp)


def gemm_tf32x3(a_hi, a_lo, b_hi, b_lo, bias, n_out, act="none", out=None):
    y = (a_hi + a_lo) @ (b_hi + b_lo).t()
    if bias is not None:
        y = y + bias
    y = _act(y, act)
    if out is not None:
        out.copy_(y)
        return out
    return y


def lstm_cell_tf32x3(x_hi, x_lo, h_hi, h_lo, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out=None):
    from se_b200 import packing
    m, kx = x_hi.shape
    hd = h_hi.shape[1]
    a = torch.cat([x_hi + x_lo, h_hi + h_lo], dim=1)
    g = a @ (w_hi + w_lo).t() + bias                       # columns: (tile j, half, gate, 16 units)
    g = g.view(m, hd // 16, 4, 16)
    i, f, gg, o = g[:, :, 0], g[:, :, 1], g[:, :, 2], g[:, :, 3]
    c = torch.sigmoid(f) * c_state.view(m, hd // 16, 16) + torch.sigmoid(i) * torch.tanh(gg)
    h = (torch.sigmoid(o) * torch.tanh(c)).reshape(m, hd)
    c_state.copy_(c.reshape(m, hd))
    hi, lo = packing.split_tf32(h)
    h_hi_out.copy_(hi)
    h_lo_out.copy_(lo)
    if h_out is not None:
        h_out.copy_(h)


def lstm_cell_tf32x3_ex(x_pair, h_pair, w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out=None):
    m, hd = c_state.shape
    if h_pair is None:
        h_pair = (torch.zeros(m, hd, dtype=c_state.dtype), torch.zeros(m, hd, dtype=c_state.dtype))
        c_state.zero_()
    lstm_cell_tf32x3(x_pair[0], x_pair[1], h_pair[0], h_pair[1], w_hi, w_lo, bias, c_state, h_hi_out, h_lo_out, h_out)


def cmul(x, m):
    return torch.view_as_real(torch.view_as_complex(x.contiguous()) * torch.view_as_complex(m.contiguous()))


def fsn_clip_inv_mean(x, strides, B, T, F, denom, wgt=None, extra=None):
    sb, st, sf = strides
    v = torch.as_strided(x, (B, T, F), (sb, st, sf))
    s = (v * wgt).sum(dim=(1, 2)) if wgt is not None else v.sum(dim=(1, 2))
    if extra is not None:
        s = s + extra.reshape(B, -1).sum(dim=1)
    return 1.0 / (s / denom + 1e-5)


def fsn_fb_input(x, strides, B, T, Tp, F, inv):
    sb, st, sf = strides
    v = torch.as_strided(x, (B, T, F), (sb, st, sf))
    mag_tm = torch.zeros(B, Tp, F)
    mag_tm[:, :T] = v
    return mag_tm, mag_tm * inv[:, None, None]


def fsn_sb_assemble(mag_tm, fb, nn, inv):
    from se_b200 import packing
    B, Tp, F = mag_tm.shape
    idx = torch.arange(F)[:, None] + torch.arange(2 * nn + 1)[None, :] - nn
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx >= F, 2 * (F - 1) - idx, idx)
    unf = mag_tm[:, :, idx]                                # [B,Tp,F,31]
    x = torch.cat([unf, fb[..., None]], dim=-1) * inv[:, None, None, None]
    x = x.permute(1, 0, 2, 3).reshape(Tp, B * F, 2 * nn + 2).contiguous()
    return packing.split_tf32(x)


def fsn_sb_fc(h, W, bias, out):
    out.copy_(h @ W.t() + bias)
    return out


# ---- fp16-pair mirrors (se_split_f16 / se_gemm_f16x3 / se_lstm_cell_f16x3 / se_conv_f16x3) ---------------------
F16_ACT_SCALE_LOG2 = 4


def _from16(pair, scale_log2):
    return (pair[0].float() + pair[1].float()) * (2.0 ** -scale_log2)


def split_f16(x2d, kpad=None, scale_log2=F16_ACT_SCALE_LOG2):
    from se_b200 import packing
    rows, k = x2d.shape
    kpad = kpad or (k + 7) // 8 * 8
    xp = x2d if kpad == k else torch.cat([x2d, x2d.new_zeros(rows, kpad - k)], dim=1)
    hi, lo, _ = packing.split_f16(xp, scale_log2)
    return hi, lo


def gemm_f16x3(a_pair, b_pair, b_scale_log2, bias, n_out, act="none", out=None, a_scale_log2=F16_ACT_SCALE_LOG2,
               act_param=0.0, alpha=1.0, res=None, want_out=True, pair_out=False, pair16_out=False,
               c16_scale_log2=F16_ACT_SCALE_LOG2):
    from se_b200 import packing
    a = _from16(a_pair, a_scale_log2)
    w = _from16(b_pair, b_scale_log2)[:, :a.shape[1]]
    y = a @ w.t()
    if bias is not None:
        y = y + bias
    y = _act(y, act, act_param) * alpha
    if res is not None:
        y = y + res
    if out is not None:
        out.copy_(y)
        y = out
    if not (pair_out or pair16_out):
        return y
    return (y, packing.split_tf32(y.contiguous()) if pair_out else None,
            packing.split_f16(y, c16_scale_log2)[:2] if pair16_out else None)


def lstm_cell_f16x3(x_pair, h_pair, cell, c_state, h_hi_out, h_lo_out, h_out=None, a_scale_log2=F16_ACT_SCALE_LOG2):
    from se_b200 import packing
    m, hd = c_state.shape
    x = _from16(x_pair, a_scale_log2)
    w = _from16((cell["w_hi"], cell["w_lo"]), cell["w_scale_log2"])
    g = x @ w[:, :x.shape[1]].t() + cell["bias"]
    if h_pair is not None:
        g = g + _from16(h_pair, a_scale_log2) @ w[:, cell["kx_pad"]:].t()
    else:
        c_state.zero_()
    g = g.view(m, hd // 16, 4, 16)
    i, f, gg, o = g[:, :, 0], g[:, :, 1], g[:, :, 2], g[:, :, 3]
    c = torch.sigmoid(f) * c_state.view(m, hd // 16, 16) + torch.sigmoid(i) * torch.tanh(gg)
    h = (torch.sigmoid(o) * torch.tanh(c)).reshape(m, hd)
    c_state.copy_(c.reshape(m, hd))
    hi, lo, _ = packing.split_f16(h, a_scale_log2)
    h_hi_out.copy_(hi)
    h_lo_out.copy_(lo)
    if h_out is not None:
        h_out.copy_(h)


def fsn_sb_assemble_f16(mag_tm, fb, nn, inv, scale_log2=F16_ACT_SCALE_LOG2):
    from se_b200 import packing
    hi, lo = fsn_sb_assemble(mag_tm, fb, nn, inv)
    h16, l16, _ = packing.split_f16(hi + lo, scale_log2)
    return h16, l16


def conv_f16x3(src0, src1, B, T, Fin, Fout, taps, sf, w_hi, w_lo, w_scale_log2, bias, Cout, act, dstF, dst_f0=0,
               dst_fstep=1, act_param=0.0, out=None, out_pair=None, out_pair16=None, glu=None,
               a_scale_log2=F16_ACT_SCALE_LOG2, out16_scale_log2=F16_ACT_SCALE_LOG2, fout1=None):
    from se_b200 import packing
    assert glu is None
    if fout1 is not None:        # two parity classes in one launch: columns [class 0 | class 1], class 1 one column further
        co = Cout // 2
        for cls, fo in ((0, Fout), (1, fout1)):
            if fo > 0:
                conv_f16x3(src0, src1, B, T, Fin, fo, taps, sf, w_hi[cls * co:(cls + 1) * co].contiguous(),
                           w_lo[cls * co:(cls + 1) * co].contiguous(), w_scale_log2, bias, co, act, dstF, dst_f0 + cls,
                           dst_fstep, act_param, out, out_pair, out_pair16, None, a_scale_log2, out16_scale_log2)
        return
    x0 = _from16(src0, a_scale_log2)
    x1 = _from16(src1, a_scale_log2) if src1 is not None else None
    c0 = x0.shape[-1]
    c1 = x1.shape[-1] if x1 is not None else 0
    p0, p1 = (c0 + 63) // 64 * 64, (c1 + 63) // 64 * 64
    wp = _from16((w_hi, w_lo), w_scale_log2).view(Cout, len(taps), p0 + p1)
    assert float(wp[:, :, c0:p0].abs().max() if p0 > c0 else 0.0) == 0.0      # the padding columns must be zero
    w = torch.cat([wp[:, :, :c0], wp[:, :, p0:p0 + c1]], dim=2).reshape(Cout, len(taps) * (c0 + c1)).t().contiguous()
    tmp = out if out is not None else torch.zeros(B, T, dstF, Cout, dtype=x0.dtype)
    if out is None and out_pair16 is not None:
        tmp.copy_(_from16(out_pair16, out16_scale_log2))
    conv_gemm(x0, x1, B, T, Fin, Fout, taps, sf, w, bias, Cout, act, tmp, dstF, dst_f0, dst_fstep, -1, None, act_param)
    if out_pair16 is not None:
        hi, lo, _ = packing.split_f16(tmp, out16_scale_log2)
        out_pair16[0].copy_(hi)
        out_pair16[1].copy_(lo)
    if out_pair is not None:
        hi, lo = packing.split_tf32(tmp)
        out_pair[0].copy_(hi)
        out_pair[1].copy_(lo)


def install(ops_module, monkeypatch):
    for name in ("conv_gemm", "linear", "conv_in1", "deconv_out1", "lstm_seq", "split_tf32", "pad_split_tf32", "gemm_tf32x3",
                 "lstm_cell_tf32x3", "fsn_clip_inv_mean", "fsn_fb_input", "fsn_sb_assemble", "fsn_sb_fc", "dccrn_mask", "conv_tf32x3", "fill_column", "lstm_seq_multi",
                 "split_f16", "gemm_f16x3", "lstm_cell_f16x3", "fsn_sb_assemble_f16", "conv_f16x3"):
        monkeypatch.setattr(ops_module, name, globals()[name])


# ---- Uformer glue mirrors -----------------------------------------------------------------------------
UF_EPS = torch.finfo(torch.float32).eps


def gemm_tf32x3_ex(a_pair, b_hi, b_lo, bias, n_out, act="none", act_param=0.0, alpha=1.0, res=None, want_f32=True,
                   want_pair=False):
    from se_b200 import packing
    y = (a_pair[0] + a_pair[1]) @ (b_hi + b_lo).t()
    if bias is not None:
        y = y + bias
    y = _act(y, act, act_param) * alpha
    if res is not None:
        y = y + res
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


def uf_prep(x):
    re, im = x[..., 0], x[..., 1]
    mag = torch.sqrt(torch.clamp(re ** 2 + im ** 2, UF_EPS))
    ph = torch.atan2(im + UF_EPS, re)
    cin = torch.stack([mag * torch.cos(ph), mag * torch.sin(ph)], -1)[:, :, 1:].contiguous()
    return mag, ph, cin, mag[:, :, 1:, None].contiguous()


def uf_fusion(c, m):
    ch = m.shape[-1]
    cr, ci = c[..., :ch], c[..., ch:]
    cm = torch.sqrt(torch.clamp(cr ** 2 + ci ** 2, UF_EPS))
    sg = torch.sigmoid(m)
    return torch.cat([cr + sg, ci + sg], -1), m + torch.sigmoid(cm)


def uf_fusion_ex(c, m, c_f32=True, c_pair=False, m_f32=True, m_pair=False):
    from se_b200 import packing
    co, mo = uf_fusion(c, m)
    return (co if c_f32 else None, packing.split_tf32(co.float()) if c_pair else None,
            mo if m_f32 else None, packing.split_tf32(mo.float()) if m_pair else None)


def group_layernorm(x, groups, gamma, beta, gate=None, post="none", slope=0.0, res=None, want_f32=True,
                    want_pair=False, eps=1e-5, out_index=None):
    from se_b200 import packing
    v = x if gate is None else x * torch.sigmoid(gate)
    shp = v.shape
    c = shp[-1] // groups
    y = F.layer_norm(v.reshape(-1, groups, c), (c,), gamma, beta, eps).reshape(shp)
    if post == "prelu":
        y = torch.where(y >= 0, y, slope * y)
    elif post == "swish":
        y = y * torch.sigmoid(y)
    if res is not None:
        y = y + res
    if out_index is not None:
        z = torch.empty_like(y).reshape(-1, groups, c)
        z[:, :, out_index.long()] = y.reshape(-1, groups, c)
        y = z.reshape(shp)
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


def attention(qkv, nheads, head_out, head_sign, nout, L, lstride, n_outer, outer_stride, n_inner, inner_stride,
              scale=0.25):
    r = qkv.shape[0]
    out = torch.zeros(r, nout * 16, dtype=qkv.dtype)
    for o in range(n_outer):
        for i in range(n_inner):
            rows = o * outer_stride + i * inner_stride + torch.arange(L) * lstride
            blk = qkv[rows]
            for h in range(nheads):
                q, k, v = blk[:, h * 48:h * 48 + 16], blk[:, h * 48 + 16:h * 48 + 32], blk[:, h * 48 + 32:h * 48 + 48]
                a = torch.softmax((q @ k.t()) * scale, dim=-1) @ v
                out[rows, head_out[h] * 16:(head_out[h] + 1) * 16] += head_sign[h] * a
    return out


def uf_mask(cmask, mdec, mag, phase):
    b, t, f = mag.shape
    mr, mi = cmask[..., 0], cmask[..., 1]
    mm = torch.sqrt(torch.clamp(mr ** 2 + mi ** 2, UF_EPS))
    rp, ip = mr / (mm + UF_EPS), mi / (mm + UF_EPS)
    mmag = F.pad(torch.tanh(mm + UF_EPS), [1, 0])
    mph = F.pad(torch.atan2(ip + UF_EPS, rp), [1, 0])
    msig = F.pad(torch.sigmoid(mdec[..., 0]), [1, 0])
    em = 0.5 * (mmag * mag + msig * mag)
    ph = phase + mph
    return torch.stack([em * torch.cos(ph), em * torch.sin(ph)], -1)


def glu_affine_act(x, scale, shift, act="elu", act_param=0.0, want_f32=True, want_pair=False):
    from se_b200 import packing
    c = x.shape[-1] // 2
    y = x[..., :c] * torch.sigmoid(x[..., c:])
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    y = _act(y, act, act_param)
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


def unary(x, act, act_param=0.0, want_f32=True, want_pair=False):
    from se_b200 import packing
    y = _act(x, act, act_param)
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


_UF_NAMES = ("glu_affine_act", "unary", "cmul", "lstm_cell_tf32x3_ex", "gemm_tf32x3_ex", "uf_prep", "uf_fusion", "uf_fusion_ex", "group_layernorm", "attention", "uf_mask")
_orig_install = install


def install(ops_module, monkeypatch):   # noqa: F811
    _orig_install(ops_module, monkeypatch)
    for name in _UF_NAMES:
        monkeypatch.setattr(ops_module, name, globals()[name])


# ---- TCM-family norm mirrors (csrc/norm.cu) ---------------------------------------------------------------
def _norm_pre(x, B, rows, C, pre, pre_slope):
    x = x.reshape(B, rows, -1)
    cin = x.shape[-1]
    if pre in ("glu", "glu_prelu"):
        v = x[..., :C] * torch.sigmoid(x[..., C:])
    elif pre == "prelu":
        v = x.repeat(1, 1, C // cin)
    else:
        v = x
    if pre in ("prelu", "glu_prelu"):
        v = torch.where(v >= 0, v, pre_slope * v)
    return v


def chan_stats(x, B, rows, C, pre="none", pre_slope=None, eps=1e-5):
    v = _norm_pre(x, B, rows, C, pre, pre_slope).double()
    m = v.mean(1)
    var = (v * v).mean(1) - m * m
    return m.float(), (1.0 / torch.sqrt(var.clamp_min(0) + eps)).float()


def cum_stats(x, B, T, F, C, pre="none", pre_slope=None, eps=1e-5, groups=1):
    v = _norm_pre(x, B, T * F, C, pre, pre_slope).double().reshape(B, T, F, groups, C // groups)
    cs, css = torch.cumsum(v.sum((2, 4)), 1), torch.cumsum((v * v).sum((2, 4)), 1)          # [B,T,G]
    cnt = (torch.arange(1, T + 1, dtype=torch.float64) * (F * C // groups))[None, :, None]
    m = cs / cnt
    var = css / cnt - m * m
    return m.float(), (1.0 / torch.sqrt(var.clamp_min(0) + eps)).float()


def chan_norm(x, B, rows, C, mean, rstd, gamma, beta, pre="none", pre_slope=None, cumulative=False, rows_per_t=1,
              post="none", post_slope=None, fir_w=None, fir_groups=1, want_f32=True, want_pair=False, stat_groups=1):
    from se_b200 import packing
    v = _norm_pre(x, B, rows, C, pre, pre_slope)
    if cumulative:
        m = mean.repeat_interleave(rows_per_t, dim=1).repeat_interleave(C // stat_groups, dim=2)
        r = rstd.repeat_interleave(rows_per_t, dim=1).repeat_interleave(C // stat_groups, dim=2)
    else:
        m, r = mean[:, None, :], rstd[:, None, :]
    y = (v - m) * r * gamma + beta
    if post == "prelu":
        y = torch.where(y >= 0, y, post_slope * y)
    elif post == "fir":
        k = fir_w.shape[1]
        w = fir_w.repeat_interleave(C // fir_groups, dim=0)          # [C, K]
        yp = F.pad(y.transpose(1, 2), (k - 1, 0))                    # [B, C, rows + K - 1]
        y = F.conv1d(yp, w[:, None, :], None, groups=C).transpose(1, 2).contiguous()
    return (y if want_f32 else None), (packing.split_tf32(y.contiguous()) if want_pair else None)


def add(a, b, want_f32=True, want_pair=False):
    from se_b200 import packing
    y = a + b
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


def axpby(a, b, ca, cb, want_f32=True, want_pair=False):
    from se_b200 import packing
    y = a + b if ca == 1 and cb == 1 else ca * a + cb * b
    return (y if want_f32 else None), (packing.split_tf32(y) if want_pair else None)


def taylor_zero(x_ri, gain, ld, want_pair=True):
    from se_b200 import packing
    f = gain.shape[-1]
    mag = torch.sqrt(x_ri[..., 0] ** 2 + x_ri[..., 1] ** 2)
    ph = torch.atan2(x_ri[..., 1], x_ri[..., 0])
    zm = gain * mag
    out = torch.zeros(gain.numel() // f, ld, dtype=gain.dtype)
    out[:, :f] = (zm * torch.cos(ph)).reshape(-1, f)
    out[:, f:2 * f] = (zm * torch.sin(ph)).reshape(-1, f)
    return out, (packing.split_tf32(out) if want_pair else None)


def gaf_update(x_re, x_im, xs_r, xs_f, gain, resi, rows, f, ld, im_off, want_f32=True, want_pair=True):
    from se_b200 import packing
    re = torch.as_strided(x_re, (rows, f), (xs_r, xs_f))
    im = torch.as_strided(x_im, (rows, f), (xs_r, xs_f))
    out = torch.zeros(rows, ld, dtype=re.dtype)
    if gain is None:
        out[:, :f], out[:, im_off:im_off + f] = re, im
    else:
        mag, ph = torch.sqrt(re * re + im * im), torch.atan2(im, re)
        xm = mag * gain.reshape(rows, f)
        out[:, :f] = xm * torch.cos(ph) + resi[:, :f]
        out[:, im_off:im_off + f] = xm * torch.sin(ph) + resi[:, im_off:im_off + f]
    return (out if want_f32 else None), (packing.split_tf32(out) if want_pair else None)


def cts_glue1(x_ri, est_mag):
    ph = torch.atan2(x_ri[..., 1], x_ri[..., 0])
    return torch.stack([x_ri[..., 0], x_ri[..., 1], est_mag * torch.cos(ph), est_mag * torch.sin(ph)], -1)


def cts_glue2(out_r, out_i, s2_in):
    return torch.stack([out_r + s2_in[..., 2], out_i + s2_in[..., 3]], -1)


_NORM_NAMES = ("chan_stats", "cum_stats", "chan_norm", "add", "axpby", "taylor_zero", "gaf_update", "cts_glue1", "cts_glue2")
_orig_install2 = install


def install(ops_module, monkeypatch):   # noqa: F811
    _orig_install2(ops_module, monkeypatch)
    for name in _NORM_NAMES:
        monkeypatch.setattr(ops_module, name, globals()[name])
