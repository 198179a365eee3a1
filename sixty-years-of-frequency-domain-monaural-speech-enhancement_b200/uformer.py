"""Drop-in ``Uformer`` (reference: Uformer/uformer.py:30-287, dilated_dualpath_conformer.py:23-78,
ff_*.py, t_att_*.py, f_att_*.py, dsconv2d_*.py, conv2d_*.py, linear_*.py, fusion.py).

Same constructor and ``forward(inputs, src)`` contract (waveforms in, ``(enhanced wav, src, est
spectrum [B,2,257,T], src spectrum)`` out; the decode scripts keep element 0,
Uformer/uformer_decode.py:45) and the same 668-entry state-dict namespace, so the shipped
checkpoints load unchanged (the ``stft.K / stft.w / istft.K / istft.w`` conv-STFT buffers are accepted
and ignored exactly as the reference's forward ignores them).

Mapping onto the engines (channels-last; complex tensors carry (re C | im C) channels):
  * complex / real U-Net convs k(5,2) s(2,1): implicit GEMM (tensor cores from 32 channels up),
    complex = stacked real with block weights, BatchNorm3d/2d folded, PReLU in the epilogue;
    ``out[..., :T]`` truncation = causal taps; decoder skip ``cat([skip, out])`` = two pointers.
  * every Linear / 1x1 conv of the conformer: tcgen05 3xTF32 GEMM with PReLU / residual epilogues;
    the 8 x 3 Real_Linear(128,16) projections of the complex attention are ONE [256 -> 384] GEMM.
  * dilated 3x3 gated convs (dilation 1..128 on T): 9-tap implicit GEMM, zero fill = "same" padding.
  * LayerNorm over channels (per real / imaginary part), gating, swish, residuals: se_group_layernorm.
  * attention over T (L = T, non-causal) and over F (L = 4): se_attention with signed head combination.
  * cross-branch ``fusion`` after every stage: se_uf_fusion.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn as nn

from . import conv_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .dccrn import _stack_complex
from .param_tree import build_param_tree

_KEYS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "uformer_keys.json")
_BUFFER_TAILS = ("running_mean", "running_var", "stft.K", "stft.w", "istft.K", "istft.w")

KN = [1, 8, 16, 32, 64, 128, 128]          # uformer.py:45 (complex and magnitude branches alike)
DIL = [1, 2, 4, 8, 16, 32, 64, 128]        # dilated_dualpath_conformer.py:39

ENC_TAPS = [(kt - 1, kf - 2) for kf in range(5) for kt in range(2)]           # in[t-1+kt, 2f+kf-2]
DEC_EVEN = [(-kt, 1 - kf // 2) for kf in (0, 2, 4) for kt in range(2)]        # f'=2m  : f = m+1-kf/2, t-kt
DEC_ODD = [(-kt, (3 - kf) // 2) for kf in (1, 3) for kt in range(2)]          # f'=2m+1: f = m+(3-kf)/2
# complex attention head table (t_att_cplx.py:58-67): (q, k, v) parts (0 = real, 1 = imag), output, sign
CPLX_HEADS = [((0, 0, 0), 0, +1.0), ((0, 1, 1), 0, -1.0), ((1, 0, 1), 0, -1.0), ((1, 1, 0), 0, -1.0),
              ((0, 0, 1), 1, +1.0), ((0, 1, 0), 1, +1.0), ((1, 0, 0), 1, +1.0), ((1, 1, 1), 1, -1.0)]


def _spec():
    # the reference's own key list (data, not code): names, shapes and order of the 668 state-dict entries
    rows = []
    for key, shape, dtype in json.load(open(_KEYS)):
        if key.endswith("num_batches_tracked"):
            kind = "counter"
        elif key.endswith(_BUFFER_TAILS):
            kind = "buffer"
        else:
            kind = "param"
        rows.append((key, tuple(shape), kind))
    return rows


def _stack_linear(wr, wi, br, bi):
    """Complex_Linear (linear_cplx.py:20-26) as one real [2out, 2in] matrix on (re | im) vectors."""
    w = torch.cat([torch.cat([wr, -wi], 1), torch.cat([wi, wr], 1)], 0)
    return w.contiguous(), torch.cat([br - bi, br + bi]).contiguous()


class _Lin:
    """[N, K] weight for the tensor-core GEMM (+ K-major copy for the FMA fallback when K % 32 != 0)."""
    __slots__ = ("hi", "lo", "kn", "bias", "n", "k")

    def __init__(self, w, bias):
        self.n, self.k = w.shape
        self.hi, self.lo = packing.split_tf32(w.contiguous())
        self.kn = packing.pad_cols(w.t().contiguous())
        self.bias = bias.contiguous() if bias is not None else None


class Uformer(nn.Module):
    def __init__(self, win_len=400, win_inc=160, fft_len=512, win_type='hanning', fid=None):
        super().__init__()
        if (win_len, win_inc, fft_len) != (400, 160, 512):
            raise NotImplementedError("se_b200 Uformer covers the shipped geometry 400/160/512 (uformer.py:33-35)")
        self.win_len, self.win_inc, self.fft_len = win_len, win_inc, fft_len
        build_param_tree(self, _spec())
        self._packed = None
        self._packed_key = None

    # ------------------------------------------------------------------------------------------------
    def _state_key(self):
        p = next(self.parameters())
        return (p.device, tuple(int(t._version) for t in self.state_dict().values()))

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items() if v.is_floating_point()}
        P = {}
        bn = lambda pre: packing.bn_fold(sd[pre + ".weight"], sd[pre + ".bias"], sd[pre + ".running_mean"],   # noqa
                                         sd[pre + ".running_var"])
        all_taps = [(kf, kt) for kf in range(5) for kt in range(2)]
        for i in range(6):
            # complex encoder level: [2Ci -> 2Co], BN3d scale shared by the real and imaginary halves
            pre = f"encoder.{i}"
            s, o = bn(pre + ".1")
            s2, o2 = torch.cat([s, s]), torch.cat([o, o])
            w = _stack_complex(sd[pre + ".0.real_conv.weight"], sd[pre + ".0.imag_conv.weight"], all_taps, False)
            br, bi = sd[pre + ".0.real_conv.bias"], sd[pre + ".0.imag_conv.bias"]
            bias = torch.cat([br - bi, br + bi]) * s2 + o2
            P[f"enc_c{i}"] = (ConvWeights(w * s2[None, :], 2 * KN[i + 1]), bias.contiguous(),
                              float(sd[pre + ".2.weight"].item()))
            pre = f"encoder_real.{i}"
            s, o = bn(pre + ".1")
            wr = sd[pre + ".0.conv.weight"]                                   # [Co, Ci, 5, 2]
            w = torch.cat([wr[:, :, kf, kt].t() for kf, kt in all_taps], 0)   # [(tap, ci), co]
            P[f"enc_m{i}"] = (ConvWeights(w * s[None, :], KN[i + 1]), (sd[pre + ".0.conv.bias"] * s + o).contiguous(),
                              float(sd[pre + ".2.weight"].item()))
        for di, idx in enumerate(range(6, 0, -1)):
            co = KN[idx - 1]
            last = idx == 1
            for branch in ("c", "m"):
                pre = f"decoder.{di}" if branch == "c" else f"decoder_real.{di}"
                if not last:
                    s, o = bn(pre + ".1")
                    slope = float(sd[pre + ".2.weight"].item())
                else:
                    ref_t = sd["conformer.ln_conformer_mag.weight"]
                    s, o, slope = ref_t.new_ones(co), ref_t.new_zeros(co), 0.0
                if branch == "c":
                    wr_, wi_ = sd[pre + ".0.real_conv.weight"], sd[pre + ".0.imag_conv.weight"]   # [2C(cat), Co, 5, 2]
                    half = wr_.shape[0] // 2                   # cat([skip, out]): first half = skip
                    s2, o2 = torch.cat([s, s]), torch.cat([o, o])
                    br, bi = sd[pre + ".0.real_conv.bias"], sd[pre + ".0.imag_conv.bias"]
                    bias = (torch.cat([br - bi, br + bi]) * s2 + o2).contiguous()

                    def parity(kfs):
                        out = []
                        for kf in kfs:
                            for kt in range(2):
                                out += [_stack_complex(wr_[:half], wi_[:half], [(kf, kt)], True),
                                        _stack_complex(wr_[half:], wi_[half:], [(kf, kt)], True)]
                        return ConvWeights(torch.cat(out, 0) * s2[None, :], 2 * co)
                else:
                    w_ = sd[pre + ".0.conv.weight"]                                             # [2C(cat), Co, 5, 2]
                    half = w_.shape[0] // 2
                    bias = (sd[pre + ".0.conv.bias"] * s + o).contiguous()

                    def parity(kfs):
                        out = []
                        for kf in kfs:
                            for kt in range(2):
                                out += [w_[:half, :, kf, kt], w_[half:, :, kf, kt]]
                        return ConvWeights(torch.cat(out, 0) * s[None, :], co)
                P[f"dec_{branch}{di}"] = (parity((0, 2, 4)), parity((1, 3)), bias, slope)
        # ---- conformer ----
        c = "conformer."

        def ln(pre):
            return sd[pre + ".weight"].contiguous(), sd[pre + ".bias"].contiguous()

        def clin(pre):
            return _Lin(*_stack_linear(sd[pre + ".real_linear.weight"], sd[pre + ".imag_linear.weight"],
                                       sd[pre + ".real_linear.bias"], sd[pre + ".imag_linear.bias"]))

        def rlin(pre):
            return _Lin(sd[pre + ".linear.weight"], sd[pre + ".linear.bias"])

        for name in ("ff1", "ff2"):
            pre = c + name + "_cplx"
            P[name + "_c"] = (ln(pre + ".layernorm_linear"), clin(pre + ".linear1"), clin(pre + ".linear2"),
                              float(sd[pre + ".prelu.weight"].item()))
            pre = c + name + "_mag"
            P[name + "_m"] = (ln(pre + ".layernorm_linear"), rlin(pre + ".linear1"), rlin(pre + ".linear2"),
                              float(sd[pre + ".prelu.weight"].item()))
        for kind, letter in (("tatt", "T"), ("fatt", "F")):
            pre = c + f"cplx_{kind}"
            h0 = pre + ".attn_heads.0"
            blocks, biases = [], []
            for hi_, ((qp, kp, vp), _, _) in enumerate(CPLX_HEADS):
                att = f"{h0}.{letter}_att{hi_ + 1}"
                for proj, part in (("query", qp), ("key", kp), ("value", vp)):
                    w = sd[f"{att}.{proj}.linear.weight"]                       # [16, 128]
                    z = torch.zeros_like(w)
                    blocks.append(torch.cat([w, z], 1) if part == 0 else torch.cat([z, w], 1))
                    biases.append(sd[f"{att}.{proj}.linear.bias"])
            P[kind + "_c"] = (ln(h0 + ".layernorm1"), _Lin(torch.cat(blocks, 0), torch.cat(biases)),
                              ln(h0 + ".layernorm2"), clin(pre + ".transform_linear"), ln(pre + ".layernorm3"),
                              float(sd[pre + ".prelu.weight"].item()))
            pre = c + f"mag_{kind}"
            h0 = pre + ".attn_heads.0"
            att = f"{h0}.{letter}_att"
            w = torch.cat([sd[f"{att}.{p_}.linear.weight"] for p_ in ("query", "key", "value")], 0)
            b_ = torch.cat([sd[f"{att}.{p_}.linear.bias"] for p_ in ("query", "key", "value")])
            P[kind + "_m"] = (ln(h0 + ".layernorm1"), _Lin(w, b_), ln(h0 + ".layernorm2"),
                              rlin(pre + ".transform_linear"), ln(pre + ".layernorm3"),
                              float(sd[pre + ".prelu.weight"].item()))
        taps9 = [(kf, kt) for kf in range(3) for kt in range(3)]
        for i in range(8):
            pre = c + f"dsconv_cplx.{i}"

            def cconv(name, taps):
                w = _stack_complex(sd[f"{pre}.{name}.real_conv.weight"], sd[f"{pre}.{name}.imag_conv.weight"], taps, False)
                br, bi = sd[f"{pre}.{name}.real_conv.bias"], sd[f"{pre}.{name}.imag_conv.bias"]
                return w, torch.cat([br - bi, br + bi]).contiguous()
            w1, b1 = cconv("conv1x1", [(0, 0)])
            wd1, bd1 = cconv("dconv1", taps9)
            wd2, bd2 = cconv("dconv2", taps9)
            ws, bs = cconv("sconv", [(0, 0)])
            P[f"ds_c{i}"] = (ln(pre + ".layernorm_conv1"), _Lin(w1.t().contiguous(), b1),
                             float(sd[pre + ".prelu.weight"].item()), (ConvWeights(wd1, 64), bd1),
                             (ConvWeights(wd2, 64), bd2), ln(pre + ".layernorm_conv2"), _Lin(ws.t().contiguous(), bs))
            pre = c + f"dsconv_real.{i}"

            def rconv(name, taps):
                w_ = sd[f"{pre}.{name}.conv.weight"]
                return torch.cat([w_[:, :, kf, kt].t() for kf, kt in taps], 0), sd[f"{pre}.{name}.conv.bias"].contiguous()
            w1, b1 = rconv("conv1x1", [(0, 0)])
            wd1, bd1 = rconv("dconv1", taps9)
            wd2, bd2 = rconv("dconv2", taps9)
            ws, bs = rconv("sconv", [(0, 0)])
            P[f"ds_m{i}"] = (ln(pre + ".layernorm_conv1"), _Lin(w1.t().contiguous(), b1),
                             float(sd[pre + ".prelu.weight"].item()), (ConvWeights(wd1, 32), bd1),
                             (ConvWeights(wd2, 32), bd2), ln(pre + ".layernorm_conv2"), _Lin(ws.t().contiguous(), bs))
        P["ln_c"] = ln(c + "ln_conformer_cplx")
        P["ln_m"] = ln(c + "ln_conformer_mag")
        self._packed = P

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _lin(x_pair, x_f32, lin: _Lin, **kw):
        """Linear on rows: tensor cores when K % 32 == 0, else the fp32 FMA kernel (K = 16 transform of the
        magnitude attention).  Returns (f32 or None, pair or None) per want_f32 / want_pair."""
        if lin.k % 32 == 0:
            return ops.gemm_tf32x3_ex(x_pair, lin.hi, lin.lo, lin.bias, lin.n, **kw)
        assert kw.get("res") is None and not kw.get("want_pair", False)
        return ops.linear(x_f32, lin.kn, lin.bias, lin.n, act=kw.get("act", "none")), None

    def _ff(self, x, groups, P_):
        """FF_Cplx / FF_Real (ff_cplx.py:21-33): LN -> linear1 -> PReLU -> linear2; y*0.5 + x."""
        (g, b), l1, l2, slope = P_
        shp = x.shape
        x2 = x.view(-1, shp[-1])
        _, y = ops.group_layernorm(x2, groups, g, b, want_f32=False, want_pair=True)
        _, h = ops.gemm_tf32x3_ex(y, l1.hi, l1.lo, l1.bias, l1.n, act="prelu", act_param=slope, want_f32=False,
                                  want_pair=True)
        out, _ = ops.gemm_tf32x3_ex(h, l2.hi, l2.lo, l2.bias, l2.n, alpha=0.5, res=x2)
        return out.view(shp)

    def _att(self, x, groups, P_, over_t, b, t, f):
        """Multihead_Attention_{T,F}_Branch[_real] (t_att_cplx.py:72-96, f_att_cplx.py:65-88)."""
        (g1, b1), qkv_lin, (g2, b2), tr, (g3, b3), slope = P_
        shp = x.shape
        x2 = x.view(-1, shp[-1])
        _, y = ops.group_layernorm(x2, groups, g1, b1, want_f32=False, want_pair=True)
        qkv, _ = ops.gemm_tf32x3_ex(y, qkv_lin.hi, qkv_lin.lo, qkv_lin.bias, qkv_lin.n)
        if groups == 2:
            heads_out = [o for _, o, _ in CPLX_HEADS]
            heads_sign = [s for _, _, s in CPLX_HEADS]
            nheads, nout = 8, 2
        else:
            heads_out, heads_sign, nheads, nout = [0], [1.0], 1, 1
        if over_t:   # sequences (b, f) over t: rows (b*T + t)*F + f
            a = ops.attention(qkv, nheads, heads_out, heads_sign, nout, t, f, b, t * f, f, 1)
        else:        # sequences (b, t) over f
            a = ops.attention(qkv, nheads, heads_out, heads_sign, nout, f, 1, b * t, f, 1, 0)
        a32, apair = ops.group_layernorm(a, groups, g2, b2, want_f32=(tr.k % 32 != 0), want_pair=(tr.k % 32 == 0))
        o, _ = self._lin(apair, a32, tr)
        out, _ = ops.group_layernorm(o, groups, g3, b3, post="prelu", slope=slope, res=x2)
        return out.view(shp)

    def _dsconv(self, x, groups, P_, dil1, dil2, b, t, f):
        """DSConv2d / DSConv2d_Real (dsconv2d_cplx.py:44-60)."""
        (g1, b1), c1, slope, (wd1, bd1), (wd2, bd2), (g2, b2), sc = P_
        shp = x.shape
        dev = x.device
        x2 = x.view(-1, shp[-1])
        _, y = ops.group_layernorm(x2, groups, g1, b1, want_f32=False, want_pair=True)
        hc = c1.n
        h32, hp = ops.gemm_tf32x3_ex(y, c1.hi, c1.lo, c1.bias, hc, act="prelu", act_param=slope,
                                     want_f32=not conv_engine.tc_eligible(hc, 0, hc, f, 1),
                                     want_pair=conv_engine.tc_eligible(hc, 0, hc, f, 1))
        hact = Act(h32.view(b, t, f, hc) if h32 is not None else None,
                   (hp[0].view(b, t, f, hc), hp[1].view(b, t, f, hc)) if hp is not None else None)
        ys = []
        for (w, bias), d in ((wd1, bd1), dil1), ((wd2, bd2), dil2):
            taps = [((kt - 1) * d, kf - 1) for kf in range(3) for kt in range(3)]
            out = conv_engine.new_act(b, t, f, hc, dev, want_f32=True, want_pair=False)
            conv_engine.conv(hact, None, b, t, f, f, taps, 1, w, bias, "none", out, f)
            ys.append(out.f32.view(-1, hc))
        _, z = ops.group_layernorm(ys[0], groups, g2, b2, gate=ys[1], post="swish", want_f32=False, want_pair=True)
        out, _ = ops.gemm_tf32x3_ex(z, sc.hi, sc.lo, sc.bias, sc.n, res=x2)
        return out.view(shp)

    def _conformer(self, c, m, b, t, f, taps=None):
        P = self._packed

        def tap(name, cc, mm):
            if taps is not None:
                taps[name + "_c"], taps[name + "_m"] = cc, mm
        c, m = self._ff(c, 2, P["ff1_c"]), self._ff(m, 1, P["ff1_m"])
        tap("ff1", c, m)
        c, m = ops.uf_fusion(c, m)
        c, m = self._att(c, 2, P["tatt_c"], True, b, t, f), self._att(m, 1, P["tatt_m"], True, b, t, f)
        tap("tatt", c, m)
        c, m = ops.uf_fusion(c, m)
        c, m = self._att(c, 2, P["fatt_c"], False, b, t, f), self._att(m, 1, P["fatt_m"], False, b, t, f)
        tap("fatt", c, m)
        c, m = ops.uf_fusion(c, m)
        for i in range(8):
            c = self._dsconv(c, 2, P[f"ds_c{i}"], DIL[i], DIL[7 - i], b, t, f)
            m = self._dsconv(m, 1, P[f"ds_m{i}"], DIL[i], DIL[7 - i], b, t, f)
            tap(f"ds{i}", c, m)
            c, m = ops.uf_fusion(c, m)
        c, m = self._ff(c, 2, P["ff2_c"]), self._ff(m, 1, P["ff2_m"])
        tap("ff2", c, m)
        c, m = ops.uf_fusion(c, m)
        c, _ = ops.group_layernorm(c, 2, *P["ln_c"])
        m, _ = ops.group_layernorm(m, 1, *P["ln_m"])
        tap("conf", c, m)
        return c, m

    def _network(self, x, taps=None):
        """x [B,T,257,2] noisy spectrum -> est [B,T,257,2]."""
        self._ensure_packed()
        P = self._packed
        b, t, fb, _ = x.shape
        dev = x.device
        mag, phase, c, m = ops.uf_prep(x)                       # c [B,T,256,2], m [B,T,256,1]
        c, m = Act(c), Act(m)
        fin = 256
        enc_c, enc_m = [], []
        tc = conv_engine.tc_eligible

        def fuse(oc, om, forms):
            """Cross-branch fusion whose results leave in the form(s) their consumers read: fp32 for the FMA convs and the
            conformer, the TF32 pair for tensor-core convs (no separate split pass).  forms = ((c_f32, c_pair),
            (m_f32, m_pair))."""
            (cf, cp), (mf, mp) = forms
            rcf, rcp, rmf, rmp = ops.uf_fusion_ex(oc, om, c_f32=cf, c_pair=cp, m_f32=mf, m_pair=mp)
            return Act(rcf, rcp), Act(rmf, rmp)

        for i in range(6):
            fo = fin // 2
            outs = []
            for src, key, ci, co in ((c, f"enc_c{i}", 2 * KN[i], 2 * KN[i + 1]), (m, f"enc_m{i}", KN[i], KN[i + 1])):
                w, bias, slope = P[key]
                out = conv_engine.new_act(b, t, fo, co, dev, want_f32=True, want_pair=False)
                conv_engine.conv(src, None, b, t, fin, fo, ENC_TAPS, 2, w, bias, "prelu", out, fo, act_param=slope)
                outs.append(out.f32)
            if taps is not None:
                taps[f"encraw{i}_c"], taps[f"encraw{i}_m"] = outs
            forms = []
            for ch, nxt, dec_co in ((2 * KN[i + 1], 2 * KN[i + 2] if i < 5 else 0, 2 * KN[i]),
                                    (KN[i + 1], KN[i + 2] if i < 5 else 0, KN[i])):
                # consumers: the next encoder conv (or the conformer after the last level) and decoder level 5 - i, which
                # takes this level as its skip source (C0 = C1 = ch, Cout = dec_co, Fout = fo)
                users_tc = [tc(ch, ch, dec_co, fo, 1)]
                if i < 5:
                    users_tc.append(tc(ch, 0, nxt, fo // 2, 2))
                want_pair = any(users_tc)
                want_f32 = (not all(users_tc)) or i == 5 or taps is not None
                forms.append((want_f32, want_pair))
            c, m = fuse(outs[0], outs[1], forms)
            enc_c.append(c)
            enc_m.append(m)
            fin = fo
        c, m = self._conformer(c.get_f32(), m.get_f32(), b, t, fin, taps)
        c, m = Act(c), Act(m)
        for di in range(6):
            fo = 2 * fin
            last = di == 5
            outs = []
            for skip, cur, key, co in ((enc_c[5 - di], c, f"dec_c{di}", 2 * KN[5 - di]),
                                       (enc_m[5 - di], m, f"dec_m{di}", KN[5 - di])):
                we, wo, bias, slope = P[key]
                act = "none" if last else "prelu"
                out = conv_engine.new_act(b, t, fo, co, dev, want_f32=True, want_pair=False)
                s0, s1 = skip, cur
                conv_engine.conv(s0, s1, b, t, fin, fin, DEC_EVEN, 1, we, bias, act, out, fo, dst_f0=0, dst_fstep=2,
                                 act_param=slope)
                conv_engine.conv(s0, s1, b, t, fin, fin, DEC_ODD, 1, wo, bias, act, out, fo, dst_f0=1, dst_fstep=2,
                                 act_param=slope)
                outs.append(out.f32)
            if taps is not None:
                taps[f"decraw{di}_c"], taps[f"decraw{di}_m"] = outs
            if last:
                c, m = fuse(outs[0], outs[1], ((True, False), (True, False)))
            else:
                forms = []
                for ch, nco in ((2 * KN[5 - di], 2 * KN[4 - di]), (KN[5 - di], KN[4 - di])):
                    nxt_tc = tc(ch, ch, nco, fo, 1)                       # the next decoder conv: C0 = skip, C1 = this
                    forms.append((not nxt_tc, nxt_tc))
                c, m = fuse(outs[0], outs[1], forms)
            fin = fo
        return ops.uf_mask(c.get_f32(), m.get_f32(), mag, phase)

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, inputs, src=None, scale=None, out_scale=None):
        """inputs [B,N] waveform (uformer.py:172).  Returns (enhanced [B, hop*(T-1)], src, est spectrum
        [B,2,257,T], None): the decode scripts use element 0 only (uformer_decode.py:45); the reference's
        ``istft(stft(src))`` round trip and src spectrum (uformer.py:186-195) are not recomputed."""
        if not inputs.is_cuda:
            raise RuntimeError("Uformer (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        from ._lib import ISTFT_SPEC
        wav = inputs.contiguous().float()
        b, n = wav.shape
        hop, win, nfft = self.win_inc, self.win_len, self.fft_len
        t, f = 1 + n // hop, nfft // 2 + 1
        x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)
        ops.stft(wav, scale, nfft, win, hop, re=x[..., 0], im=x[..., 1])
        est = self._network(x)
        length = hop * (t - 1)
        out = torch.empty(b, length, device=wav.device, dtype=torch.float32)
        ops.istft(ISTFT_SPEC, est[..., 0], est[..., 1], None, None, nfft, win, hop, out, length, out_scale=out_scale)
        return out, src, est.permute(0, 3, 2, 1), None
