"""Oracle networks: functional torch-CPU restatements of the reference forward() bodies.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function takes the
reference's own state-dict (same keys as the shipped ``BEST_MODEL/*.pth``) and an
input tensor, and computes what the reference ``nn.Module`` computes in
``eval()`` mode, with the same ATen CPU operators the reference would dispatch to
(``conv2d``, ``conv_transpose2d``, ``batch_norm``, ``lstm`` ...), so that it is
also a fair CPU timing arm.  Pinned against the unmodified reference modules by
``oracle/make_golden.py`` (max-abs difference recorded in the fixture).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm default, used by every BN in CRN.py / LSTM.py


def _bn(x, sd, prefix):
    """eval-mode BatchNorm{1,2}d: running statistics, affine."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, BN_EPS)


def lstm(x, sd, prefix, num_layers):
    """nn.LSTM(batch_first=True), zero initial state, eval.  x [B,T,I] -> [B,T,H]."""
    flat = []
    for l in range(num_layers):
        flat += [sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"],
                 sd[f"{prefix}.bias_ih_l{l}"], sd[f"{prefix}.bias_hh_l{l}"]]
    hid = flat[1].shape[1]
    b = x.shape[0]
    h0 = x.new_zeros(num_layers, b, hid)
    c0 = x.new_zeros(num_layers, b, hid)
    out, _, _ = torch._VF.lstm(x, (h0, c0), flat, True, num_layers, 0.0, False, False, True)
    return out


def lstm_manual(x, sd, prefix, num_layers):
    """Same as :func:`lstm` but spelled out step by step (gate order i,f,g,o) -- used by the
    tests to show that the ATen fused op and the textbook recurrence agree."""
    b, t, _ = x.shape
    for l in range(num_layers):
        w_ih, w_hh = sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"]
        bias = sd[f"{prefix}.bias_ih_l{l}"] + sd[f"{prefix}.bias_hh_l{l}"]
        hid = w_hh.shape[1]
        h = x.new_zeros(b, hid)
        c = x.new_zeros(b, hid)
        outs = []
        xp = x @ w_ih.t() + bias
        for s in range(t):
            g = xp[:, s] + h @ w_hh.t()
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        x = torch.stack(outs, dim=1)
    return x


# ----------------------------------------------------------------------------------------
# CRN  (CRN/CRN.py)
# ----------------------------------------------------------------------------------------
def crn_forward(sd, x, taps=None):
    """crn_net.forward, CRN/CRN.py:23-33.  x [B,T,161] magnitude -> [B,T,161] estimated magnitude.

    Encoder (CRN.py:35-71): 5 x [pad one frame on top, Conv2d k(2,3) s(1,2), BN, ELU].
    LSTM (CRN.py:20,27-31): [B,T,256*4] (channel-major flatten), 2 layers of 1024.
    Decoder (CRN.py:73-109): 5 x [cat skip, ConvTranspose2d k(2,3) s(1,2), (de4: left-pad F by 1,
    CRN.py:92-97), drop last frame, BN, ELU]; de5 ends in BN(1)+Softplus.
    ``taps`` (optional dict) receives per-layer activations for bisecting.
    """
    x = x.unsqueeze(1)
    b, _, t, _ = x.shape
    skips = []
    for i in range(5):
        x = F.pad(x, (0, 0, 1, 0))
        x = F.conv2d(x, sd[f"en.en_module.{i}.1.weight"], sd[f"en.en_module.{i}.1.bias"], stride=(1, 2))
        x = F.elu(_bn(x, sd, f"en.en_module.{i}.2"))
        skips.append(x)
        if taps is not None:
            taps[f"en{i + 1}"] = x
    x = x.permute(0, 2, 1, 3).contiguous().view(b, t, -1)
    x = lstm(x, sd, "lstm", 2)
    if taps is not None:
        taps["lstm"] = x
    x = x.view(b, t, 256, 4).permute(0, 2, 1, 3).contiguous()
    for i in range(5):
        x = torch.cat((x, skips[-(i + 1)]), dim=1)
        x = F.conv_transpose2d(x, sd[f"de.de_module.{i}.0.weight"], sd[f"de.de_module.{i}.0.bias"],
                               stride=(1, 2))
        if i == 3:
            x = F.pad(x, (1, 0, 0, 0))
        x = x[:, :, :-1, :]
        bn_idx = 3 if i == 3 else 2
        x = _bn(x, sd, f"de.de_module.{i}.{bn_idx}")
        x = F.softplus(x) if i == 4 else F.elu(x)
        if taps is not None:
            taps[f"de{i + 1}"] = x
    return x.squeeze(1)


# ----------------------------------------------------------------------------------------
# LSTM  (LSTM/LSTM.py)
# ----------------------------------------------------------------------------------------
def lstm_net_forward(sd, x, taps=None):
    """lstm_net.forward, LSTM/LSTM.py:24-29.  x [B,T,161] -> [B,T,161].

    BatchNorm1d over the 161 bins (LSTM.py:16,25), LSTM 161->1024 (:17), LSTM 1024->1024 x2 (:18),
    Linear 1024->161 + Softplus (:19-22).
    """
    x = _bn(x.permute(0, 2, 1).contiguous(), sd, "bn").permute(0, 2, 1).contiguous()
    x = lstm(x, sd, "lstm1", 1)
    if taps is not None:
        taps["lstm1"] = x
    x = lstm(x, sd, "lstm2", 2)
    if taps is not None:
        taps["lstm2"] = x
    x = F.softplus(F.linear(x, sd["fc.0.weight"], sd["fc.0.bias"]))
    return x


# ----------------------------------------------------------------------------------------
# FullSubNet  (FullSubNet/fullsubnet_net_sa/model.py)
# ----------------------------------------------------------------------------------------
def _fsn_unfold(x, num_neighbor):
    """BaseModel.unfold, base_model.py:12-42.  x [B,1,F,T] -> [B,F,2n+1,T] (reflect pad on F)."""
    b, _, f, t = x.shape
    if num_neighbor < 1:
        return x.permute(0, 2, 1, 3).reshape(b, f, 1, t)
    p = F.pad(x, [0, 0, num_neighbor, num_neighbor], mode="reflect")        # [B,1,F+2n,T]
    idx = torch.arange(f)[:, None] + torch.arange(2 * num_neighbor + 1)[None, :]   # [F, 2n+1]
    return p[:, 0][:, idx]                                                   # [B,F,2n+1,T]


def _fsn_norm(x):
    """offline_laplace_norm, base_model.py:196-209: divide by the utterance mean (+1e-5)."""
    mu = torch.mean(x, dim=(1, 2, 3), keepdim=True)
    return x / (mu + 1e-5)


def _fsn_sequence(sd, prefix, x, act):
    """SequenceModel.forward, sequence_model.py:66-84.  x [N,F,T] -> [N,Fo,T]."""
    o = lstm(x.permute(0, 2, 1).contiguous(), sd, prefix + ".sequence_model", 2)
    o = F.linear(o, sd[prefix + ".fc_output_layer.weight"], sd[prefix + ".fc_output_layer.bias"])
    if act == "relu":
        o = F.relu(o)
    return o.permute(0, 2, 1).contiguous()


def fullsubnet_forward(sd, noisy_mag, look_ahead=2, sb_num_neighbors=15, fb_num_neighbors=0, taps=None):
    """Model.forward, model.py:68-118, with PER-UTTERANCE semantics: the reference decodes one
    file at a time (B=1), where the ``if batch_size > 1: drop_band`` branch (model.py:101-104)
    never runs; a batched oracle must therefore skip it.
    noisy_mag [B,1,257,T] -> complex mask [B,2,257,T]."""
    x = F.pad(noisy_mag, [0, look_ahead])                                    # model.py:79
    b, c, f, t = x.shape
    fb_in = _fsn_norm(x).reshape(b, c * f, t)                                # :84
    fb_out = _fsn_sequence(sd, "fb_model", fb_in, "relu").reshape(b, 1, f, t)   # :85
    if taps is not None:
        taps["fb_out"] = fb_out
    fb_unf = _fsn_unfold(fb_out, fb_num_neighbors)                           # :88-89
    nm_unf = _fsn_unfold(x, sb_num_neighbors)                                # :92-93
    sb_in = _fsn_norm(torch.cat([nm_unf, fb_unf], dim=2))                    # :96-97
    if taps is not None:
        taps["sb_in"] = sb_in
    sb_in = sb_in.reshape(b * f, sb_in.shape[2], t)                          # :106-110
    mask = _fsn_sequence(sd, "sb_model", sb_in, None)                        # :113
    mask = mask.reshape(b, f, 2, t).permute(0, 2, 1, 3).contiguous()         # :114
    return mask[:, :, :, look_ahead:]                                        # :117


# ----------------------------------------------------------------------------------------
# DCCRN  (DCCRN/DCCRN_cprs.py + un-vendored complexnn, restated in oracle/complexnn_restated.py)
# ----------------------------------------------------------------------------------------
def _cconv(x, sd, pre, transpose):
    """ComplexConv2d / ComplexConvTranspose2d k(5,2) s(2,1): DCCRN_cprs.py:66-72, 108-115."""
    r, i = torch.chunk(x, 2, 1)
    wr, br, wi, bi = (sd[pre + ".real_conv.weight"], sd[pre + ".real_conv.bias"], sd[pre + ".imag_conv.weight"],
                      sd[pre + ".imag_conv.bias"])
    if transpose:
        f = lambda v, w, b: F.conv_transpose2d(v, w, b, stride=(2, 1), padding=(2, 0), output_padding=(1, 0))  # noqa
    else:
        f = lambda v, w, b: F.conv2d(v, w, b, stride=(2, 1), padding=(2, 0))  # noqa: E731
    return torch.cat([f(r, wr, br) - f(i, wi, bi), f(r, wi, bi) + f(i, wr, br)], 1)


def _lstm1(x, sd, pre):
    """single-layer nn.LSTM, batch_first=False.  x [T,B,I]."""
    return lstm(x.transpose(0, 1), {f"l.weight_ih_l0": sd[pre + ".weight_ih_l0"], "l.weight_hh_l0": sd[pre + ".weight_hh_l0"],
                                    "l.bias_ih_l0": sd[pre + ".bias_ih_l0"], "l.bias_hh_l0": sd[pre + ".bias_hh_l0"]},
                "l", 1).transpose(0, 1)


def dccrn_forward(sd, inputs, crop_first=True, taps=None, masking_mode='E'):
    """DCCRN.forward with use_clstm=True (DCCRN_cprs.py:142-226); masking_mode 'E' is what every script uses, 'C' and
    'R' are the other two branches (:221-224).
    inputs [B,2,257,T] compressed RI -> [B,2,257,T].  ``crop_first``: decoder keeps ``out[...,1:]``
    (DCCRN_cprs.py:199); the DCCRN_SNR variant keeps ``[..., :-1]`` (DCCRN_SNR/DCCRN.py:159)."""
    spec_mags = torch.norm(inputs, dim=1)                                   # :164
    spec_phase = torch.atan2(inputs[:, -1], inputs[:, 0])                   # :165
    out = inputs[:, :, 1:]                                                  # :166  drop DC
    enc = []
    for idx in range(6):
        out = F.pad(out, [1, 0, 0, 0])                                      # causal pad (complexnn)
        out = _cconv(out, sd, f"encoder.{idx}.0", False)
        out = _bn(out, sd, f"encoder.{idx}.1")
        out = F.prelu(out, sd[f"encoder.{idx}.2.weight"])
        enc.append(out)
        if taps is not None:
            taps[f"enc{idx}"] = out
    b, c, d, t = out.shape
    out = out.permute(3, 0, 1, 2)                                           # :175
    r = out[:, :, :c // 2].reshape(t, b, c // 2 * d)
    i = out[:, :, c // 2:].reshape(t, b, c // 2 * d)
    for l in range(2):                                                      # :182 NavieComplexLSTM x2
        pre = f"enhance.{l}"
        r2r, r2i = _lstm1(r, sd, pre + ".real_lstm"), _lstm1(r, sd, pre + ".imag_lstm")
        i2r, i2i = _lstm1(i, sd, pre + ".real_lstm"), _lstm1(i, sd, pre + ".imag_lstm")
        r, i = r2r - i2i, i2r + r2i
        if pre + ".r_trans.weight" in sd:
            r = F.linear(r, sd[pre + ".r_trans.weight"], sd[pre + ".r_trans.bias"])
            i = F.linear(i, sd[pre + ".i_trans.weight"], sd[pre + ".i_trans.bias"])
    if taps is not None:
        taps["rnn_r"], taps["rnn_i"] = r, i
    out = torch.cat([r.reshape(t, b, c // 2, d), i.reshape(t, b, c // 2, d)], 2).permute(1, 2, 3, 0)   # :183-194
    for idx in range(6):
        e = enc[-1 - idx]
        orr, oi = torch.chunk(out, 2, 1)
        er, ei = torch.chunk(e, 2, 1)
        out = torch.cat([orr, er, oi, ei], 1)                               # complex_cat :197
        out = _cconv(out, sd, f"decoder.{idx}.0", True)
        if idx < 5:
            out = _bn(out, sd, f"decoder.{idx}.1")
            out = F.prelu(out, sd[f"decoder.{idx}.2.weight"])
        out = out[..., 1:] if crop_first else out[..., :-1]                 # :199
        if taps is not None:
            taps[f"dec{idx}"] = out
    mask_real = F.pad(out[:, 0], [0, 0, 1, 0])                              # :203-204
    mask_imag = F.pad(out[:, 1], [0, 0, 1, 0])
    if masking_mode != 'E':
        real, imag = inputs[:, 0], inputs[:, -1]                            # :163
        if masking_mode == 'C':                                             # :221-222
            return torch.stack([real * mask_real - imag * mask_imag, real * mask_imag + imag * mask_real], 1)
        return torch.stack([real * mask_real, imag * mask_imag], 1)         # :223-224 'R'
    mask_mags = (mask_real ** 2 + mask_imag ** 2) ** 0.5                    # :207
    real_phase = mask_real / (mask_mags + 1e-8)
    imag_phase = mask_imag / (mask_mags + 1e-8)
    mask_phase = torch.atan2(imag_phase, real_phase)
    est_mags = torch.tanh(mask_mags) * spec_mags                            # :216-217
    est_phase = spec_phase + mask_phase
    return torch.stack([est_mags * torch.cos(est_phase), est_mags * torch.sin(est_phase)], 1)


# ----------------------------------------------------------------------------------------
# Uformer  (Uformer/uformer.py and the block files it imports)
# ----------------------------------------------------------------------------------------
UF_EPS = torch.finfo(torch.float32).eps          # EPSILON in every Uformer file
_UF_KN = [1, 8, 16, 32, 64, 128, 128]            # uformer.py:45
_UF_DIL = [1, 2, 4, 8, 16, 32, 64, 128]          # dilated_dualpath_conformer.py:39


def _uf_fusion(c, m):
    """fusion.py:13-19.  c [N,C,F,T,2], m [N,C,F,T]."""
    cm = torch.sqrt(torch.clamp(c[..., 0] ** 2 + c[..., 1] ** 2, UF_EPS))
    sg = torch.sigmoid(m)
    return torch.stack([c[..., 0] + sg, c[..., 1] + sg], -1), m + torch.sigmoid(cm)


def _uf_cconv(x, sd, pre, transpose, **kw):
    """ComplexConv2d_Encoder / _Decoder (conv2d_cplx.py:31-38, 62-68): real/imag cross terms, then
    truncation to the input length on T."""
    f = F.conv_transpose2d if transpose else F.conv2d
    wr, br, wi, bi = (sd[pre + ".real_conv.weight"], sd[pre + ".real_conv.bias"], sd[pre + ".imag_conv.weight"],
                      sd[pre + ".imag_conv.bias"])
    xr, xi = x[..., 0], x[..., 1]
    t = xr.shape[-1]
    orr = f(xr, wr, br, **kw) - f(xi, wi, bi, **kw)
    oi = f(xi, wr, br, **kw) + f(xr, wi, bi, **kw)
    return torch.stack([orr[..., :t], oi[..., :t]], -1)


def _uf_rconv(x, sd, pre, transpose, **kw):
    """RealConv2d_Encoder / _Decoder (conv2d_real.py:33-36, 57-61)."""
    f = F.conv_transpose2d if transpose else F.conv2d
    return f(x, sd[pre + ".conv.weight"], sd[pre + ".conv.bias"], **kw)[..., :x.shape[-1]]


def _uf_ln(x, sd, pre, dim):
    """nn.LayerNorm over channel axis ``dim`` (the x.transpose(dim, -1) idiom, e.g. ff_cplx.py:23)."""
    return F.layer_norm(x.transpose(dim, -1), (x.shape[dim],), sd[pre + ".weight"], sd[pre + ".bias"]).transpose(dim, -1)


def _uf_clinear(x, sd, pre):
    """Complex_Linear (linear_cplx.py:20-26) on [..., C, 2]-last-but-one layout [..., in, 2]."""
    lr = lambda v: F.linear(v, sd[pre + ".real_linear.weight"], sd[pre + ".real_linear.bias"])   # noqa: E731
    li = lambda v: F.linear(v, sd[pre + ".imag_linear.weight"], sd[pre + ".imag_linear.bias"])   # noqa: E731
    xr, xi = x[..., 0], x[..., 1]
    return torch.stack([lr(xr) - li(xi), lr(xi) + li(xr)], -1)


def _uf_rlinear(x, sd, pre):
    return F.linear(x, sd[pre + ".linear.weight"], sd[pre + ".linear.bias"])


def _uf_ff(x, sd, pre, cplx):
    """FF_Cplx / FF_Real (ff_cplx.py:21-33, ff_real.py:21-33)."""
    y = _uf_ln(x, sd, pre + ".layernorm_linear", 1)
    if cplx:
        y = y.transpose(1, 3)                                   # N T F C 2
        y = _uf_clinear(y, sd, pre + ".linear1")
        y = F.prelu(y, sd[pre + ".prelu.weight"])
        y = _uf_clinear(y, sd, pre + ".linear2").transpose(1, 3)
    else:
        y = y.transpose(1, 3)
        y = _uf_rlinear(y, sd, pre + ".linear1")
        y = F.prelu(y, sd[pre + ".prelu.weight"])
        y = _uf_rlinear(y, sd, pre + ".linear2").transpose(1, 3)
    return y * 0.5 + x


def _uf_att1(q, k, v, sd, pre):
    """T_att / F_att (t_att_cplx.py:21-38): softmax(q k^T / sqrt(16)) v with three Real_Linear(128,16)."""
    qq, kk, vv = _uf_rlinear(q, sd, pre + ".query"), _uf_rlinear(k, sd, pre + ".key"), _uf_rlinear(v, sd, pre + ".value")
    e = torch.softmax(qq @ kk.transpose(1, 2) / 16 ** 0.5, dim=-1)
    return e @ vv


def _uf_attention(x, sd, pre, cplx, over_t):
    """Multihead_Attention_{T,F}_Branch[_real] (t_att_cplx.py:40-96, f_att_cplx.py:33-88, *_real.py)."""
    letter = "T" if over_t else "F"
    h0 = pre + ".attn_heads.0"
    if cplx:
        n, c, f, t, ri = x.shape
        s = x.permute(0, 2, 3, 1, 4) if over_t else x.permute(0, 3, 2, 1, 4)      # N F T C 2 | N T F C 2
        s = s.contiguous().view(-1, s.shape[2], c, ri)
        s = F.layer_norm(s.transpose(2, 3), (c,), sd[h0 + ".layernorm1.weight"], sd[h0 + ".layernorm1.bias"]).transpose(2, 3)
        re, im = s[..., 0], s[..., 1]
        a = lambda i, q, k, v: _uf_att1(q, k, v, sd, f"{h0}.{letter}_att{i}")   # noqa: E731
        real_att = a(1, re, re, re) - a(2, re, im, im) - a(3, im, re, im) - a(4, im, im, re)
        imag_att = a(5, re, re, im) + a(6, re, im, re) + a(7, im, re, re) - a(8, im, im, im)
        o = torch.stack([real_att, imag_att], -1)
        o = F.layer_norm(o.transpose(2, 3), (16,), sd[h0 + ".layernorm2.weight"], sd[h0 + ".layernorm2.bias"]).transpose(2, 3)
        o = _uf_clinear(o, sd, pre + ".transform_linear")
        o = o.contiguous().view(n, f, t, c, ri).permute(0, 3, 1, 2, 4) if over_t else \
            o.contiguous().view(n, t, f, c, ri).permute(0, 3, 2, 1, 4)
        o = F.prelu(_uf_ln(o, sd, pre + ".layernorm3", 1), sd[pre + ".prelu.weight"])
        return o + x
    n, c, f, t = x.shape
    s = x.permute(0, 2, 3, 1) if over_t else x.permute(0, 3, 2, 1)
    s = s.contiguous().view(-1, s.shape[2], c)
    o = F.layer_norm(s, (c,), sd[h0 + ".layernorm1.weight"], sd[h0 + ".layernorm1.bias"])
    o = _uf_att1(o, o, o, sd, f"{h0}.{letter}_att")
    o = F.layer_norm(o, (16,), sd[h0 + ".layernorm2.weight"], sd[h0 + ".layernorm2.bias"])
    o = _uf_rlinear(o, sd, pre + ".transform_linear")
    o = o.contiguous().view(n, f, t, c) if over_t else o.contiguous().view(n, t, f, c)
    o = F.prelu(F.layer_norm(o, (c,), sd[pre + ".layernorm3.weight"], sd[pre + ".layernorm3.bias"]),
                sd[pre + ".prelu.weight"])
    o = o.permute(0, 3, 1, 2) if over_t else o.permute(0, 3, 2, 1)
    return o + x


def _uf_dsconv(x, sd, pre, cplx, d1, d2):
    """DSConv2d / DSConv2d_Real (dsconv2d_cplx.py:44-60)."""
    conv = _uf_cconv if cplx else _uf_rconv
    y = _uf_ln(x, sd, pre + ".layernorm_conv1", 1)
    y = conv(y, sd, pre + ".conv1x1", False)
    y = F.prelu(y, sd[pre + ".prelu.weight"])
    y1 = conv(y, sd, pre + ".dconv1", False, padding=(1, d1), dilation=(1, d1))
    y2 = conv(y, sd, pre + ".dconv2", False, padding=(1, d2), dilation=(1, d2))
    y = y1 * torch.sigmoid(y2)
    y = _uf_ln(y, sd, pre + ".layernorm_conv2", 1)
    y = y * torch.sigmoid(y)
    return x + conv(y, sd, pre + ".sconv", False)


def uformer_forward(sd, x_re, x_im, taps=None):
    """Uformer.forward from the noisy spectrum on (uformer.py:197-262); the STFT / iSTFT around it live in
    oracle.decode.  x_re, x_im [B,257,T] -> est (real, imag) [B,257,T]."""
    x_re, x_im = x_re.unsqueeze(1), x_im.unsqueeze(1)
    mag = torch.sqrt(torch.clamp(x_re ** 2 + x_im ** 2, UF_EPS))            # :197
    phase = torch.atan2(x_im + UF_EPS, x_re)
    mag0, phase0 = mag, phase
    out = torch.stack([mag * torch.cos(phase), mag * torch.sin(phase)], -1)[:, :, 1:]   # :205-209
    mag = mag[:, :, 1:]                                                      # :210
    enc_c, enc_m = [], []
    for i in range(6):                                                       # :214-219
        out = _uf_cconv(out, sd, f"encoder.{i}.0", False, stride=(2, 1), padding=(2, 1))
        out = F.batch_norm(out, sd[f"encoder.{i}.1.running_mean"], sd[f"encoder.{i}.1.running_var"],
                           sd[f"encoder.{i}.1.weight"], sd[f"encoder.{i}.1.bias"], False, 0.0, BN_EPS)
        out = F.prelu(out, sd[f"encoder.{i}.2.weight"])
        mag = _uf_rconv(mag, sd, f"encoder_real.{i}.0", False, stride=(2, 1), padding=(2, 1))
        mag = F.prelu(_bn(mag, sd, f"encoder_real.{i}.1"), sd[f"encoder_real.{i}.2.weight"])
        out, mag = _uf_fusion(out, mag)
        enc_c.append(out)
        enc_m.append(mag)
    c = "conformer."                                                         # dilated_dualpath_conformer.py:53-78
    out, mag = _uf_fusion(_uf_ff(out, sd, c + "ff1_cplx", True), _uf_ff(mag, sd, c + "ff1_mag", False))
    out, mag = _uf_fusion(_uf_attention(out, sd, c + "cplx_tatt", True, True), _uf_attention(mag, sd, c + "mag_tatt", False, True))
    out, mag = _uf_fusion(_uf_attention(out, sd, c + "cplx_fatt", True, False), _uf_attention(mag, sd, c + "mag_fatt", False, False))
    for i in range(8):
        out = _uf_dsconv(out, sd, c + f"dsconv_cplx.{i}", True, _UF_DIL[i], _UF_DIL[7 - i])
        mag = _uf_dsconv(mag, sd, c + f"dsconv_real.{i}", False, _UF_DIL[i], _UF_DIL[7 - i])
        out, mag = _uf_fusion(out, mag)
    out, mag = _uf_fusion(_uf_ff(out, sd, c + "ff2_cplx", True), _uf_ff(mag, sd, c + "ff2_mag", False))
    out, mag = _uf_ln(out, sd, c + "ln_conformer_cplx", 1), _uf_ln(mag, sd, c + "ln_conformer_mag", 1)
    if taps is not None:
        taps["conf_c"], taps["conf_m"] = out, mag
    for di in range(6):                                                      # uformer.py:225-232
        kw = dict(stride=(2, 1), padding=(2, 0), output_padding=(1, 0))
        out = _uf_cconv(torch.cat([enc_c[-1 - di], out], 1), sd, f"decoder.{di}.0", True, **kw)
        mag = _uf_rconv(torch.cat([enc_m[-1 - di], mag], 1), sd, f"decoder_real.{di}.0", True, **kw)
        if di < 5:
            out = F.batch_norm(out, sd[f"decoder.{di}.1.running_mean"], sd[f"decoder.{di}.1.running_var"],
                               sd[f"decoder.{di}.1.weight"], sd[f"decoder.{di}.1.bias"], False, 0.0, BN_EPS)
            out = F.prelu(out, sd[f"decoder.{di}.2.weight"])
            mag = F.prelu(_bn(mag, sd, f"decoder_real.{di}.1"), sd[f"decoder_real.{di}.2.weight"])
        out, mag = _uf_fusion(out, mag)
    mag = F.pad(torch.sigmoid(mag), [0, 0, 1, 0])[:, 0] * mag0[:, 0]          # :236-239
    mr, mi = out[..., 0], out[..., 1]
    mm = torch.sqrt(torch.clamp(mr ** 2 + mi ** 2, UF_EPS))                  # :244
    rp, ip = mr / (mm + UF_EPS), mi / (mm + UF_EPS)
    mmag = F.pad(torch.tanh(mm + UF_EPS), [0, 0, 1, 0])                      # :247,249
    mph = F.pad(torch.atan2(ip + UF_EPS, rp), [0, 0, 1, 0])                  # :248,250
    est_mag = (mmag[:, 0] * mag0[:, 0] + mag) * 0.5                          # :254,262
    est_ph = phase0[:, 0] + mph[:, 0]                                        # :257
    return est_mag * torch.cos(est_ph), est_mag * torch.sin(est_ph)


# ----------------------------------------------------------------------------------------
# GCRN  (GCRN/GCRN_noncprs.py)
# ----------------------------------------------------------------------------------------
def _gcrn_glu(x, sd, pre, transpose, output_padding=(0, 0)):
    """GluConv2d / GluConvTranspose2d (GCRN_noncprs.py:42-83): conv1(x) * sigmoid(conv2(x)), k(1,3) s(1,2)."""
    if transpose:
        f = lambda n: F.conv_transpose2d(x, sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"], stride=(1, 2),   # noqa: E731
                                         output_padding=output_padding)
    else:
        f = lambda n: F.conv2d(x, sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"], stride=(1, 2))   # noqa: E731
    return f("conv1") * torch.sigmoid(f("conv2"))


def gcrn_forward(sd, x, taps=None):
    """Net.forward, GCRN_noncprs.py:136-165.  x [B,2,T,161] RI -> [B,2,T,161] RI."""
    enc = []
    out = x
    for i in range(1, 6):                                                    # :138-142
        out = F.elu(_bn(_gcrn_glu(out, sd, f"conv{i}", False), sd, f"bn{i}"))
        enc.append(out)
        if taps is not None:
            taps[f"e{i}"] = out
    e5 = out
    # GLSTM.forward, :22-39
    b, c, t, f = out.shape
    o = out.transpose(1, 2).contiguous().view(b, t, -1)
    o = torch.chunk(o, 2, dim=-1)
    o = torch.stack([_lstm1(o[i].transpose(0, 1), sd, f"glstm.lstm_list1.{i}").transpose(0, 1) for i in range(2)], dim=-1)
    o = torch.flatten(o, start_dim=-2, end_dim=-1)                           # interleaves the two groups (:28-29)
    o = F.layer_norm(o, (1024,), sd["glstm.ln1.weight"], sd["glstm.ln1.bias"])
    o = torch.chunk(o, 2, dim=-1)
    o = torch.cat([_lstm1(o[i].transpose(0, 1), sd, f"glstm.lstm_list2.{i}").transpose(0, 1) for i in range(2)], dim=-1)
    o = F.layer_norm(o, (1024,), sd["glstm.ln2.weight"], sd["glstm.ln2.bias"])
    out = o.view(b, t, c, -1).transpose(1, 2).contiguous()
    if taps is not None:
        taps["glstm"] = out
    out = torch.cat((out, e5), dim=1)                                        # :147
    res = []
    for br in (1, 2):                                                        # :149-159: skip is ELU'd a second time
        d = out
        for lvl, skip in ((5, enc[3]), (4, enc[2]), (3, enc[1]), (2, enc[0])):
            op = (0, 1) if lvl == 2 else (0, 0)
            d = _bn(_gcrn_glu(d, sd, f"conv{lvl}_t_{br}", True, op), sd, f"bn{lvl}_t_{br}")
            d = F.elu(torch.cat((d, skip), dim=1))
        d = F.elu(_bn(_gcrn_glu(d, sd, f"conv1_t_{br}", True), sd, f"bn1_t_{br}"))
        if taps is not None:
            taps[f"d1_{br}"] = d
        res.append(F.linear(d, sd[f"fc{br}.weight"], sd[f"fc{br}.bias"]))     # :161-162
    return torch.cat(res, dim=1)


# ----------------------------------------------------------------------------------------
# DPCRN  (DPCRN/DPCRN.py)
# ----------------------------------------------------------------------------------------
def _lstm_dir(x, sd, prefix, layer, reverse):
    """One direction of one nn.LSTM layer, batch_first.  x [B,L,I] -> [B,L,H] (outputs at their own positions)."""
    sfx = f"_l{layer}" + ("_reverse" if reverse else "")
    one = {"l.weight_ih_l0": sd[f"{prefix}.weight_ih{sfx}"], "l.weight_hh_l0": sd[f"{prefix}.weight_hh{sfx}"],
           "l.bias_ih_l0": sd[f"{prefix}.bias_ih{sfx}"], "l.bias_hh_l0": sd[f"{prefix}.bias_hh{sfx}"]}
    if reverse:
        return lstm(x.flip(1), one, "l", 1).flip(1)
    return lstm(x, one, "l", 1)


def _bilstm(x, sd, prefix, num_layers):
    """nn.LSTM(bidirectional=True, batch_first=True): layer l+1 sees cat(forward, backward) of layer l."""
    for l in range(num_layers):
        x = torch.cat((_lstm_dir(x, sd, prefix, l, False), _lstm_dir(x, sd, prefix, l, True)), dim=-1)
    return x


def _dprnn(sd, x):
    """DPRNN.forward, DPCRN.py:60-98.  x [B,C=128,T,F=4] -> same shape."""
    b, c, t, f = x.shape
    xt = x.permute(0, 2, 3, 1).contiguous()                                   # [B,T,F,C]   (:64)
    out = _bilstm(xt.view(-1, f, c), sd, "dprnn.intra_rnn", 2)                # over F      (:70-71)
    out = F.linear(out, sd["dprnn.intra_fc.weight"], sd["dprnn.intra_fc.bias"]).view(b, -1, f, c)
    out = F.layer_norm(out, (f, c), sd["dprnn.ln1.weight"], sd["dprnn.ln1.bias"])
    intra = out + xt                                                          # :76
    out = intra.permute(0, 2, 1, 3).contiguous().view(-1, t, c)               # [B*F,T,C]   (:80-81)
    out = lstm(out, sd, "dprnn.inter_rnn", 2)
    out = F.linear(out, sd["dprnn.inter_fc.weight"], sd["dprnn.inter_fc.bias"]).view(b, -1, t, c)
    out = out.permute(0, 2, 1, 3).contiguous()                                # [B,T,F,C]   (:87-88)
    out = F.layer_norm(out, (f, c), sd["dprnn.ln2.weight"], sd["dprnn.ln2.bias"]) + intra
    return out.permute(0, 3, 1, 2).contiguous()


def dpcrn_forward(sd, inpt, taps=None):
    """dpcrn.forward, DPCRN.py:23-42.  inpt [B,2,T,161] RI -> [B,2,T,161] RI (complex ratio mask applied inside).

    Encoder (:100-138): 5 x [pad one frame on top, Conv2d k(2,3) s(1,2), BN, PReLU(1)]; the SAME DPRNN twice
    (:28-29); decoder (:140-179): 5 x [cat skip, ConvTranspose2d k(2,3) s(1,2), (de4: left-pad F by 1), drop
    the last frame, BN, PReLU]; de5 has neither BN nor activation (:166-168)."""
    x = inpt
    skips = []
    for i in range(5):
        x = F.pad(x, (0, 0, 1, 0))
        x = F.conv2d(x, sd[f"en.en_module.{i}.1.weight"], sd[f"en.en_module.{i}.1.bias"], stride=(1, 2))
        x = F.prelu(_bn(x, sd, f"en.en_module.{i}.2"), sd[f"en.en_module.{i}.3.weight"])
        skips.append(x)
        if taps is not None:
            taps[f"en{i + 1}"] = x
    x = _dprnn(sd, x)
    if taps is not None:
        taps["dp1"] = x
    x = _dprnn(sd, x)
    if taps is not None:
        taps["dp2"] = x
    for i in range(5):
        x = torch.cat((x, skips[-(i + 1)]), dim=1)
        x = F.conv_transpose2d(x, sd[f"de.de_module.{i}.0.weight"], sd[f"de.de_module.{i}.0.bias"], stride=(1, 2))
        if i == 3:
            x = F.pad(x, (1, 0, 0, 0))
        x = x[:, :, :-1, :]
        if i < 4:
            bn_idx, act_idx = (3, 4) if i == 3 else (2, 3)
            x = F.prelu(_bn(x, sd, f"de.de_module.{i}.{bn_idx}"), sd[f"de.de_module.{i}.{act_idx}.weight"])
        if taps is not None:
            taps[f"de{i + 1}"] = x
    mr, mi = x[:, 0], x[:, 1]
    xr, xi = inpt[:, 0], inpt[:, 1]
    return torch.stack((xr * mr - xi * mi, xr * mi + xi * mr), dim=1)        # :39-42


# ----------------------------------------------------------------------------------------
# CTSNet  (CTSNet/Step1_network.py, CTSNet/Step2_network.py; CTSNet_new/* swap the norms)
# ----------------------------------------------------------------------------------------
def _cts_norm(x, sd, pre, cumulative):
    """nn.InstanceNorm{1,2}d(affine=True) -- per (clip, channel) statistics over (T[,F]) even in eval()
    (track_running_stats=False; Step1_network.py:48,164) -- or, for the ``_new`` variants, the causal
    CumulativeLayerNorm{1,2}d (CTSNet_new/Step1_network.py:213-286): statistics over (C[,F]) of all frames <= t."""
    if not cumulative:
        return F.instance_norm(x, None, None, sd[pre + ".weight"], sd[pre + ".bias"], True, 0.0, 1e-5)
    red = [1, 3] if x.dim() == 4 else [1]
    step_sum = x.sum(red, keepdim=True)
    step_pow = x.pow(2).sum(red, keepdim=True)
    cum_sum, cum_pow = torch.cumsum(step_sum, dim=2), torch.cumsum(step_pow, dim=2)
    per = x.shape[1] * (x.shape[3] if x.dim() == 4 else 1)
    cnt = torch.arange(per, per * (x.shape[2] + 1), per, dtype=x.dtype).view([1, 1, -1] + [1] * (x.dim() - 3))
    mean = cum_sum / cnt
    var = (cum_pow - 2 * mean * cum_sum) / cnt + mean.pow(2)
    return (x - mean) / (var + 1e-5).sqrt() * sd[pre + ".gain"] + sd[pre + ".bias"]


def _cts_gate_conv(x, sd, pre, transpose):
    """Gate_Conv (Step1_network.py:127-151): conv(x) * sigmoid(gate_conv(x)); encoder convs pad one frame on top
    (causal), decoder transposed convs drop the last output frame (Chomp_T(1))."""
    if transpose:
        f = lambda n: F.conv_transpose2d(x, sd[f"{pre}.{n}.0.weight"], sd[f"{pre}.{n}.0.bias"],   # noqa: E731
                                         stride=(1, 2))[:, :, :-1, :]
    else:
        xp = F.pad(x, (0, 0, 1, 0))
        f = lambda n: F.conv2d(xp, sd[f"{pre}.{n}.1.weight"], sd[f"{pre}.{n}.1.bias"], stride=(1, 2))   # noqa: E731
    return f("conv") * torch.sigmoid(f("gate_conv"))


def _cts_tcm(x, sd, pre, d, branches, cumulative):
    """Glu / glu (Step1_network.py:163-193, Step2_network.py:124-156), dilation d.  x [B,256,T]."""
    def branch(u, name):
        u = F.prelu(u, sd[f"{pre}.{name}.0.weight"])
        u = _cts_norm(u, sd, f"{pre}.{name}.1", cumulative)
        w = sd[f"{pre}.{name}.2.weight"]                       # ShareSepConv: one FIR shared by all channels
        k = w.shape[-1]
        u = F.conv1d(F.pad(u, (k - 1, 0)), w.expand(u.shape[1], 1, k).contiguous(), None, groups=u.shape[1])
        return F.conv1d(F.pad(u, (4 * d, 0)), sd[f"{pre}.{name}.4.weight"], None, dilation=d)
    u = F.conv1d(x, sd[f"{pre}.in_conv.weight"])
    u = branch(u, branches[0]) * torch.sigmoid(branch(u, branches[1]))
    u = F.prelu(u, sd[f"{pre}.out_conv.0.weight"])
    u = _cts_norm(u, sd, f"{pre}.out_conv.1", cumulative)
    return F.conv1d(u, sd[f"{pre}.out_conv.2.weight"]) + x


def _cts_encoder(sd, x, pre, cumulative, taps=None):
    skips = []
    for i in range(5):
        x = _cts_gate_conv(x, sd, f"{pre}.{i}.0", False)
        x = F.prelu(_cts_norm(x, sd, f"{pre}.{i}.1", cumulative), sd[f"{pre}.{i}.2.weight"])
        skips.append(x)
        if taps is not None:
            taps[f"e{i + 1}"] = x
    return x, skips


def _cts_decoder(sd, x, skips, pre, cumulative):
    for i in range(5):
        x = torch.cat((x, skips[-(i + 1)]), dim=1)
        x = _cts_gate_conv(x, sd, f"{pre}.{i}.0", True)
        x = F.prelu(_cts_norm(x, sd, f"{pre}.{i}.1", cumulative), sd[f"{pre}.{i}.2.weight"])
    return x.squeeze(1)


def ctsnet_step1_forward(sd, x, cumulative=False, taps=None):
    """Step1_net.forward, Step1_network.py:21-41.  x [B,T,161] magnitude -> [B,T,161] magnitude."""
    x, skips = _cts_encoder(sd, x.unsqueeze(1), "en.en", cumulative, taps)
    b, _, t, _ = x.shape
    x = x.permute(0, 1, 3, 2).contiguous().view(b, -1, t)                    # [B, 64*4, T], feature = c*4 + f
    acc = torch.zeros_like(x)
    for s in (1, 2, 3):                                                      # :28-33
        for j in range(6):
            x = _cts_tcm(x, sd, f"tcm{s}.tcm_list.{j}", 2 ** j, ("left_conv", "right_conv"), cumulative)
        acc = acc + x
    if taps is not None:
        taps["tcm"] = acc
    x = acc.view(b, 64, 4, t).permute(0, 1, 3, 2).contiguous()
    x = _cts_decoder(sd, x, skips, "de.de", cumulative)
    return F.softplus(F.linear(x, sd["de.de6.0.weight"], sd["de.de6.0.bias"]))   # :113-115


def ctsnet_step2_forward(sd, inpt, X=6, R=3, cumulative=False, taps=None):
    """Step2_net.forward, Step2_network.py:23-37.  inpt [B,4,T,161] (noisy RI, stage-1 RI) -> [B,2,T,161]."""
    x, skips = _cts_encoder(sd, inpt, "en.en_module", cumulative, taps)
    b, _, t, _ = x.shape
    x = x.permute(0, 1, 3, 2).contiguous().view(b, -1, t)
    acc = torch.zeros_like(x)
    for r in range(R):
        for j in range(X):
            x = _cts_tcm(x, sd, f"tcm_list.{r}.glu_list.{j}", 2 ** j, ("ori_conv", "att_ori"), cumulative)
        acc = acc + x
    if taps is not None:
        taps["tcm"] = acc
    x = acc.view(b, 64, 4, t).permute(0, 1, 3, 2).contiguous()
    outs = []
    for br in ("de_r", "de_i"):
        d = _cts_decoder(sd, x, skips, f"{br}.de_list", cumulative)
        outs.append(F.linear(d, sd[f"{br}.de6.0.weight"], sd[f"{br}.de6.0.bias"]))
    return torch.stack(outs, dim=1)


# ----------------------------------------------------------------------------------------
# TaylorSENet  (TaylorSENet/TaylorSENet.py; TaylorSENet_new swaps InstanceNorm for cumulative LayerNorm)
# configuration of taylorsenet_decode_vb.py:11-13: cin=2, k1=(1,3), k2=(2,3), c=64, kd1=5, cd1=64, d_feat=256,
# dilations=[1,2,5,9], p=2, order_num=3, intra/inter_connect='cat', causal, no conformer, U2-Net, nothing shared.
# ----------------------------------------------------------------------------------------
TAYLOR_DILATIONS = (1, 2, 5, 9)


def _ty_gate_conv(x, sd, pre, kt, transpose):
    """GateConv2d / GateConvTranspose2d (TaylorSENet.py:549-603): ONE conv with 2*C outputs, chunk, a * sigmoid(b).
    kt > 1: causal top pad (conv, key ``conv.1``) or Chomp_T (transposed conv, key ``conv.0``)."""
    if transpose:
        key = f"{pre}.conv.0" if kt > 1 else f"{pre}.conv"
        y = F.conv_transpose2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=(1, 2))
        if kt > 1:
            y = y[:, :, :-(kt - 1), :]
    else:
        key = f"{pre}.conv.1" if kt > 1 else f"{pre}.conv"
        y = F.conv2d(F.pad(x, (0, 0, kt - 1, 0)), sd[key + ".weight"], sd[key + ".bias"], stride=(1, 2))
    a, b = y.chunk(2, dim=1)
    return a * torch.sigmoid(b)


def _ty_unet_module(x, sd, pre, kt_in, scale, transpose, cum):
    """En_unet_module (TaylorSENet.py:441-496): gated in_conv + norm + PReLU, then a small U-Net of ``scale``
    Conv2dunit / Deconv2dunit levels (k2 = (2,3), stride (1,2), intra_connect = 'cat'), residual."""
    r = _ty_gate_conv(x, sd, f"{pre}.in_conv.0", kt_in, transpose)
    r = F.prelu(_cts_norm(r, sd, f"{pre}.in_conv.1", cum), sd[f"{pre}.in_conv.2.weight"])
    x, xs = r, []
    for i in range(scale):                                                   # Conv2dunit :498-519
        x = F.conv2d(F.pad(x, (0, 0, 1, 0)), sd[f"{pre}.enco.{i}.conv.1.weight"], sd[f"{pre}.enco.{i}.conv.1.bias"],
                     stride=(1, 2))
        x = F.prelu(_cts_norm(x, sd, f"{pre}.enco.{i}.conv.2", cum), sd[f"{pre}.enco.{i}.conv.3.weight"])
        xs.append(x)
    for i in range(scale):                                                   # Deconv2dunit :521-547
        if i > 0:
            x = torch.cat((x, xs[-(i + 1)]), dim=1)
        x = F.conv_transpose2d(x, sd[f"{pre}.deco.{i}.deconv.0.weight"], sd[f"{pre}.deco.{i}.deconv.0.bias"],
                               stride=(1, 2))[:, :, :-1, :]
        x = F.prelu(_cts_norm(x, sd, f"{pre}.deco.{i}.deconv.2", cum), sd[f"{pre}.deco.{i}.deconv.3.weight"])
    return r + x


def _ty_u2_encoder(x, sd, pre, cum):
    """U2Net_Encoder (TaylorSENet.py:336-370): 4 modules (scale 4..1; the first has the (2,5) kernel) + last_conv."""
    outs = []
    for i, scale in enumerate((4, 3, 2, 1)):
        x = _ty_unet_module(x, sd, f"{pre}.meta_unet_list.{i}", 2 if i == 0 else 1, scale, False, cum)
        outs.append(x)
    x = _ty_gate_conv(x, sd, f"{pre}.last_conv.0", 1, False)
    x = F.prelu(_cts_norm(x, sd, f"{pre}.last_conv.1", cum), sd[f"{pre}.last_conv.2.weight"])
    outs.append(x)
    return x, outs


def _ty_u2_decoder(x, en, sd, pre, cum):
    """U2Net_Decoder, inter_connect='cat' (TaylorSENet.py:404-438): 4 modules (scale 1..4) + gated (2,5) deconv to 16
    channels + norm + PReLU + 1x1 conv + sigmoid -> gain [B,T,161]."""
    for i, scale in enumerate((1, 2, 3, 4)):
        x = _ty_unet_module(torch.cat((x, en[-(i + 1)]), dim=1), sd, f"{pre}.meta_unet_list.{i}", 1, scale, True, cum)
    x = torch.cat((x, en[0]), dim=1)
    x = _ty_gate_conv(x, sd, f"{pre}.last_conv.0", 2, True)
    x = F.prelu(_cts_norm(x, sd, f"{pre}.last_conv.1", cum), sd[f"{pre}.last_conv.2.weight"])
    x = torch.sigmoid(F.conv2d(x, sd[f"{pre}.last_conv.3.weight"], sd[f"{pre}.last_conv.3.bias"]))
    return x.squeeze(1)


def _ty_tcm(x, sd, pre, d, cum, kd=5):
    """SqueezedTCM (TaylorSENet.py:641-685), causal."""
    def branch(u, name):
        u = F.prelu(u, sd[f"{pre}.{name}.0.weight"])
        u = _cts_norm(u, sd, f"{pre}.{name}.1", cum)
        return F.conv1d(F.pad(u, ((kd - 1) * d, 0)), sd[f"{pre}.{name}.3.weight"], None, dilation=d)
    u = F.conv1d(x, sd[f"{pre}.in_conv.weight"])
    u = branch(u, "left_conv") * torch.sigmoid(branch(u, "right_conv"))
    u = _cts_norm(F.prelu(u, sd[f"{pre}.out_conv.0.weight"]), sd, f"{pre}.out_conv.1", cum)
    return F.conv1d(u, sd[f"{pre}.out_conv.2.weight"]) + x


def _ty_tcms(x, sd, pre, cum, p=2):
    for i in range(p):
        for j, d in enumerate(TAYLOR_DILATIONS):
            x = _ty_tcm(x, sd, f"{pre}.{i}.tcm_list.{j}", d, cum)
    return x


def taylorsenet_forward(sd, inputs, cumulative=False, order_num=3, taps=None):
    """TaylorSENet.forward, TaylorSENet.py:66-94.  inputs [B,2,T,161] (compressed) RI -> [B,2,T,161]."""
    import math
    cum = cumulative
    mag, phase = torch.norm(inputs, dim=1), torch.atan2(inputs[:, -1], inputs[:, 0])       # :73
    # ZeroOrderBlock.forward :139-153
    en_x, en_list = _ty_u2_encoder(inputs, sd, "zeroorderblock.en", cum)
    b, c, t, f = en_x.shape
    x = en_x.transpose(-2, -1).contiguous().view(b, c * f, t)
    x = _ty_tcms(x, sd, "zeroorderblock.tcms", cum)
    gain = _ty_u2_decoder(x.view(b, c, f, t).transpose(-2, -1).contiguous(), en_list, sd, "zeroorderblock.de", cum)
    if taps is not None:
        taps["gain"] = gain
    zmag = gain * mag                                                                       # :75
    zero = torch.stack((zmag * torch.cos(phase), zmag * torch.sin(phase)), dim=1)           # :76
    head, _ = _ty_u2_encoder(inputs, sd, "separate_en", cum)                                # :79
    head = head.transpose(-2, -1).contiguous().view(b, -1, t)
    if taps is not None:
        taps["head"] = head
    out, pre_term = zero, zero
    for k in range(order_num):                                                              # :84-93
        hp = f"highorderblock_list.{k}"
        x1 = pre_term.transpose(-2, -1).contiguous().view(b, -1, t)                         # HighOrderBlock.forward :191-214
        x = F.conv1d(torch.cat((head, x1), dim=1), sd[hp + ".in_conv.weight"], sd[hp + ".in_conv.bias"])
        x = _ty_tcms(x, sd, hp + ".tcms", cum)
        xr = F.conv1d(x, sd[hp + ".real_resi.weight"], sd[hp + ".real_resi.bias"]).transpose(-2, -1)
        xi = F.conv1d(x, sd[hp + ".imag_resi.weight"], sd[hp + ".imag_resi.bias"]).transpose(-2, -1)
        update = torch.stack((xr, xi), dim=1) + k * pre_term
        pre_term = update
        out = out + update / math.factorial(k + 1)
    return out


# ----------------------------------------------------------------------------------------
# G2Net  (G2Net_new/gaf_net_320.py: cumulative LayerNorm; G2Net_VB/gaf_net_320.py: InstanceNorm)
# configuration of com_decode.py:23: gaf_base(3, 64, 2, 4, 4, [1,2,5,9], 256+161*2, 256, 256, (2,3), (1,3), 64, 'cat', 3,
# is_aux=False, encoder_type='U2Net', tcm_type='full-band')
# ----------------------------------------------------------------------------------------
G2_DILAS = (1, 2, 5, 9)


def _g2_unet_module(x, sd, pre, scale, cum):
    """En_unet_module (gaf_net_320.py:384-431): gated (2,3)/(2,5) in_conv (two convs, causal) + norm + PReLU, then a
    U-Net of ``scale`` Conv2dunit / Deconv2dunit levels with k2 = (1,3) (no time kernel), intra 'cat', residual."""
    r = _cts_gate_conv(x, sd, f"{pre}.in_conv.0", False)
    r = F.prelu(_cts_norm(r, sd, f"{pre}.in_conv.1", cum), sd[f"{pre}.in_conv.2.weight"])
    x, xs = r, []
    for i in range(scale):
        x = F.conv2d(x, sd[f"{pre}.enco.{i}.conv.0.weight"], sd[f"{pre}.enco.{i}.conv.0.bias"], stride=(1, 2))
        x = F.prelu(_cts_norm(x, sd, f"{pre}.enco.{i}.conv.1", cum), sd[f"{pre}.enco.{i}.conv.2.weight"])
        xs.append(x)
    for i in range(scale):
        if i > 0:
            x = torch.cat((x, xs[-(i + 1)]), dim=1)
        x = F.conv_transpose2d(x, sd[f"{pre}.deco.{i}.deconv.0.weight"], sd[f"{pre}.deco.{i}.deconv.0.bias"], stride=(1, 2))
        x = F.prelu(_cts_norm(x, sd, f"{pre}.deco.{i}.deconv.1", cum), sd[f"{pre}.deco.{i}.deconv.2.weight"])
    return r + x


def _g2_glu(x, sd, pre, d, cum):
    """Glu (gaf_net_320.py:245-274): single-branch squeezed TCM, k = 3, causal."""
    u = F.conv1d(x, sd[f"{pre}.in_conv.weight"])
    u = _cts_norm(F.prelu(u, sd[f"{pre}.left_conv.0.weight"]), sd, f"{pre}.left_conv.1", cum)
    u = F.conv1d(F.pad(u, (2 * d, 0)), sd[f"{pre}.left_conv.3.weight"], None, dilation=d)
    u = _cts_norm(F.prelu(u, sd[f"{pre}.out_conv.0.weight"]), sd, f"{pre}.out_conv.1", cum)
    return F.conv1d(u, sd[f"{pre}.out_conv.2.weight"]) + x


def _g2_tcm_head(x, sd, pre, cum, tcm_num=2):
    """nn.Sequential(*Tcm_list x tcm_num, Conv1d(256, 161, 1)[, Sigmoid])  (:135-139, :170-177)."""
    for i in range(tcm_num):
        for j, d in enumerate(G2_DILAS):
            x = _g2_glu(x, sd, f"{pre}.{i}.tcm_list.{j}", d, cum)
    return F.conv1d(x, sd[f"{pre}.{tcm_num}.weight"], sd[f"{pre}.{tcm_num}.bias"])


def g2net_forward(sd, inpt, cumulative=True, stage_num=3, taps=None):
    """gaf_base.forward, gaf_net_320.py:73-87 (is_aux=False).  inpt [B,2,T,161] -> list of stage outputs [B,2,161,T]."""
    cum = cumulative
    b, _, t, _ = inpt.shape
    x = inpt
    for i, scale in enumerate((4, 3, 2, 1)):                                 # U2Net_Encoder :277-303
        x = _g2_unet_module(x, sd, f"en.meta_unet_list.{i}", scale, cum)
    x = _cts_gate_conv(x, sd, "en.last_conv.0", False)
    x = F.prelu(_cts_norm(x, sd, "en.last_conv.1", cum), sd["en.last_conv.2.weight"])
    feat = x.transpose(-2, -1).contiguous().view(b, -1, t)
    if taps is not None:
        taps["feat"] = feat
    pre_x = inpt.transpose(-2, -1).contiguous()                              # [B,2,F,T]
    outs = []
    for s in range(stage_num):                                               # GAF_module.forward :104-115
        g = f"gafs.{s}"
        pre_mag, pre_phase = torch.norm(pre_x, dim=1), torch.atan2(pre_x[:, -1], pre_x[:, 0])
        xin = torch.cat((feat, pre_x.view(b, -1, t)), 1)
        gb, fb = g + ".glance_branch", g + ".focus_branch"
        xg = F.conv1d(xin, sd[gb + ".in_conv_main.weight"], sd[gb + ".in_conv_main.bias"]) * \
            torch.sigmoid(F.conv1d(xin, sd[gb + ".in_conv_gate.0.weight"], sd[gb + ".in_conv_gate.0.bias"]))
        gain = torch.sigmoid(_g2_tcm_head(xg, sd, gb + ".mstcm_filter", cum))
        xf = F.conv1d(xin, sd[fb + ".in_conv_main.weight"], sd[fb + ".in_conv_main.bias"]) * \
            torch.sigmoid(F.conv1d(xin, sd[fb + ".in_conv_gate.0.weight"], sd[fb + ".in_conv_gate.0.bias"]))
        resi = torch.stack((_g2_tcm_head(xf, sd, fb + ".mstcm_r", cum), _g2_tcm_head(xf, sd, fb + ".mstcm_i", cum)), 1)
        x_mag = pre_mag * gain
        pre_x = torch.stack((x_mag * torch.cos(pre_phase), x_mag * torch.sin(pre_phase)), 1) + resi
        outs.append(pre_x)
    return outs
