"""Drop-in ``TaylorSENet`` (reference: TaylorSENet/TaylorSENet.py:8-94 and its causal-norm twin
TaylorSENet_new/TaylorSENet.py).

    TaylorSENet(cin=2, k1=(1,3), k2=(2,3), c=64, kd1=5, cd1=64, d_feat=256, dilations=[1,2,5,9], p=2, fft_num=320,
                order_num=3, intra_connect='cat', inter_connect='cat', is_causal=True, is_conformer=False, is_u2=True,
                is_param_share=False, is_encoder_share=False).forward(x [B,2,T,161]) -> [B,2,T,161]

i.e. the configuration taylorsenet_decode_vb.py:11-13 builds; the 811 state-dict keys are the reference's, so the
shipped ``BEST_MODEL/*_taylor_*_model.pth`` load unchanged.  ``cumulative=True`` selects the TaylorSENet_new
norms (CumulativeLayerNorm, ``.gain`` keys).  Inference only.

Mapping onto the kernels (channels-last activations, see ctsnet.py for the shared machinery):
  * En_unet_module (:441-496): gated in_conv (ONE conv with 2C outputs, :549-603) + norm + PReLU, then a U-Net of
    ``scale`` plain Conv2dunit / Deconv2dunit levels (k(2,3), stride (1,2), causal pad / Chomp_T folded into the tap
    tables, intra 'cat' = second source pointer), residual add (se_add).  Every conv is one implicit GEMM
    (tensor cores for C = 64) followed by the two-pass utterance norm (se_chan_stats / se_cum_stats + se_chan_norm).
  * SqueezedTCM (:641-685) = the CTSNet TCM without ShareSepConv.
  * The Taylor recursion (:84-93) lives on "RI rows" [B*T, 352] = [re(161) | im(161) | zero pad to a multiple of 32]:
    HighOrderBlock.in_conv over cat(feature_head, pre_term) is two tensor-core GEMMs (K = 256 and K = 352, the second
    adds the first in its epilogue), real_resi / imag_resi are ONE GEMM with N = 352 that writes the next RI row,
    update / out accumulate with se_axpby; the zeroth-order term gain * |X| * (cos, sin)(angle X) is se_taylor_zero.
"""
from __future__ import annotations

import json
import math
import os

import torch

from . import conv_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .ctsnet import _CtsBase
from .param_tree import build_param_tree

N_BINS = 161
RI_LD = 352                      # 2 * 161 padded to a multiple of 32 (tensor-core K)
DILATIONS = (1, 2, 5, 9)
_HERE = os.path.dirname(os.path.abspath(__file__))

ENC_F = [161, 79, 39, 19, 9, 4]
DEC23_EVEN = [(0, 0), (0, -1), (-1, 0), (-1, -1)]      # (kt,kf) = (0,0),(0,2),(1,0),(1,2)
DEC23_ODD = [(0, 0), (-1, 0)]                          # (0,1),(1,1)


def _keys(cumulative):
    with open(os.path.join(_HERE, "taylor_new_keys.json" if cumulative else "taylor_keys.json")) as f:
        return json.load(f)


class _U2Base(_CtsBase):
    """U2-Net machinery shared by TaylorSENet and G2Net: gated convs, En_unet_module, the encoder loop."""

    def _inner_taps(self):
        """Tap tables of the inner Conv2dunit / Deconv2dunit levels (k2 = (2,3) here; G2Net overrides with (1,3))."""
        return packing.CONV23_TAPS, DEC23_EVEN, DEC23_ODD

    def _plain_block(self, tmp, b, t, f, c, gamma, beta, slope, want_f32, want_pair):
        st = self._stats(tmp, b, t, f, c, "none")
        f32, pair = self._norm(tmp, b, t, f, c, st, gamma, beta, "none", post="prelu", post_slope=slope,
                               want_f32=want_f32, want_pair=want_pair)
        v = lambda z: z.view(b, t, f, c)      # noqa: E731
        return Act(v(f32) if f32 is not None else None, (v(pair[0]), v(pair[1])) if pair is not None else None)

    def _gate_conv(self, src, skip, b, t, fin, packed, transpose, kf_full):
        """Runs the (parity classes of the) gated conv into a fresh [B,T,fout,2C] fp32 tensor."""
        classes, bias = packed
        dev = (src.f32 if src.f32 is not None else src.pair[0]).device
        if not transpose:
            fout = (fin - kf_full) // 2 + 1
            w, tp = classes[0]
            tmp = Act(torch.empty(b, t, fout, w.cout, device=dev, dtype=torch.float32))
            conv_engine.conv(src, skip, b, t, fin, fout, tp, 2, w, bias, "none", tmp, fout)
            return tmp.f32, fout
        fout = 2 * (fin - 1) + kf_full
        tmp = Act(torch.empty(b, t, fout, classes[0][0].cout, device=dev, dtype=torch.float32))
        for par, (w, tp) in enumerate(classes):
            conv_engine.conv(src, skip, b, t, fin, (fout - par + 1) // 2, tp, 1, w, bias, "none", tmp, fout, dst_f0=par,
                             dst_fstep=2)
        return tmp.f32, fout

    def _module(self, src, skip, b, t, fin, name, kf_in, want_f32=False):
        """En_unet_module.forward (:480-496).  Returns (Act, fout)."""
        m = self._packed[name]
        tmp, f0 = self._gate_conv(src, skip, b, t, fin, m["in"], m["transpose"], kf_in)
        r = self._gated_block(tmp, b, t, f0, 64, *m["in_norm"], m["in_slope"], want_f32=True, want_pair=True)
        x, xs, f = r, [], f0
        dev = tmp.device
        enc_taps, dec_even, dec_odd = self._inner_taps()
        for j in range(m["scale"]):                                                   # Conv2dunit
            w, bias, gamma, beta, slope = m["enco"][j]
            f2 = (f - 3) // 2 + 1
            tmp = Act(torch.empty(b, t, f2, 64, device=dev, dtype=torch.float32))
            conv_engine.conv(x, None, b, t, f, f2, enc_taps, 2, w, bias, "none", tmp, f2)
            x = self._plain_block(tmp.f32, b, t, f2, 64, gamma, beta, slope, want_f32=False, want_pair=True)
            xs.append(x)
            f = f2
        for j in range(m["scale"]):                                                   # Deconv2dunit (+ intra 'cat')
            even, odd, bias, gamma, beta, slope = m["deco"][j]
            sk = xs[-(j + 1)] if j > 0 else None
            f2 = 2 * f + 1
            tmp = Act(torch.empty(b, t, f2, 64, device=dev, dtype=torch.float32))
            conv_engine.conv(x, sk, b, t, f, f + 1, dec_even, 1, even, bias, "none", tmp, f2, dst_f0=0, dst_fstep=2)
            conv_engine.conv(x, sk, b, t, f, f, dec_odd, 1, odd, bias, "none", tmp, f2, dst_f0=1, dst_fstep=2)
            last = j == m["scale"] - 1
            x = self._plain_block(tmp.f32, b, t, f2, 64, gamma, beta, slope, want_f32=last, want_pair=not last)
            f = f2
        assert f == f0
        f32, pair = ops.add(r.f32, x.f32, want_f32=want_f32, want_pair=True)          # x_resi + x  (:494)
        return Act(f32, pair), f0

    def _u2_encoder(self, x, b, t, name, want_last_f32, taps=None):
        h, f, outs = Act(x), N_BINS, []
        for i in range(4):
            h, f = self._module(h, None, b, t, f, f"{name}{i}", 5 if i == 0 else 3)
            outs.append(h)
        packed, (gamma, beta), slope = self._packed[f"{name}_last"]
        tmp, f = self._gate_conv(h, None, b, t, f, packed, False, 3)
        h = self._gated_block(tmp, b, t, f, 64, gamma, beta, slope, want_f32=want_last_f32, want_pair=True)
        outs.append(h)
        return outs


class TaylorSENet(_U2Base):
    def __init__(self, cin=2, k1=(1, 3), k2=(2, 3), c=64, kd1=5, cd1=64, d_feat=256, dilations=(1, 2, 5, 9), p=2,
                 fft_num=320, order_num=3, intra_connect="cat", inter_connect="cat", is_causal=True,
                 is_conformer=False, is_u2=True, is_param_share=False, is_encoder_share=False, cumulative=False):
        super().__init__(cumulative)
        cfg = (cin, tuple(k1), tuple(k2), c, kd1, cd1, d_feat, tuple(dilations), p, fft_num, order_num, intra_connect,
               inter_connect, is_causal, is_conformer, is_u2, is_param_share, is_encoder_share)
        if cfg != (2, (1, 3), (2, 3), 64, 5, 64, 256, DILATIONS, 2, 320, 3, "cat", "cat", True, False, True, False, False):
            raise NotImplementedError("se_b200.TaylorSENet implements the configuration of taylorsenet_decode_vb.py:11-13")
        self.order_num, self.p = order_num, p
        build_param_tree(self, [(k, tuple(s), "param") for k, s in _keys(self.cumulative).items()])

    # -- weight packing --------------------------------------------------------------------------
    def _pack_gate(self, sd, key, transpose, kt):
        """GateConv2d / GateConvTranspose2d -> list of (ConvWeights, taps, class) + bias; N = [a | b] as the reference's
        chunk(2, dim=1) orders it."""
        w, bias = sd[key + ".weight"], sd[key + ".bias"].contiguous()
        if not transpose:                                           # [2C, Ci, kt, kf]
            kf = w.shape[-1]
            wk = w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()
            return [(ConvWeights(wk, w.shape[0]), [(a - (kt - 1), f) for a in range(kt) for f in range(kf)])], bias
        kf = w.shape[-1]                                            # [Ci, 2C, kt, kf]; out[t, 2m + (k&1)] += in[t-a, m - k//2]
        classes = []
        for par in (0, 1):
            tk = [(a, f) for a in range(kt) for f in range(par, kf, 2)]
            wk = torch.cat([w[:, :, a, f] for a, f in tk], 0).contiguous()
            classes.append((ConvWeights(wk, w.shape[1]), [(-a, -(f // 2)) for a, f in tk]))
        return classes, bias

    def _pack_module(self, sd, P, pre, name, scale, transpose, kt_in):
        key = f"{pre}.in_conv.0.conv" + ((".0" if transpose else ".1") if kt_in > 1 else "")
        m = {"in": self._pack_gate(sd, key, transpose, kt_in), "in_norm": self._norm_params(sd, f"{pre}.in_conv.1"),
             "in_slope": sd[f"{pre}.in_conv.2.weight"].contiguous(), "scale": scale, "transpose": transpose,
             "enco": [], "deco": []}
        for j in range(scale):
            w = sd[f"{pre}.enco.{j}.conv.1.weight"]                                            # [64, 64, 2, 3]
            wk = w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()
            m["enco"].append((ConvWeights(wk, 64), sd[f"{pre}.enco.{j}.conv.1.bias"].contiguous(),
                              *self._norm_params(sd, f"{pre}.enco.{j}.conv.2"), sd[f"{pre}.enco.{j}.conv.3.weight"].contiguous()))
            w = sd[f"{pre}.deco.{j}.deconv.0.weight"]                                          # [64 | 128, 64, 2, 3]
            even = torch.cat([w[:, :, 0, 0], w[:, :, 0, 2], w[:, :, 1, 0], w[:, :, 1, 2]], 0).contiguous()
            odd = torch.cat([w[:, :, 0, 1], w[:, :, 1, 1]], 0).contiguous()
            m["deco"].append((ConvWeights(even, 64), ConvWeights(odd, 64), sd[f"{pre}.deco.{j}.deconv.0.bias"].contiguous(),
                              *self._norm_params(sd, f"{pre}.deco.{j}.deconv.2"),
                              sd[f"{pre}.deco.{j}.deconv.3.weight"].contiguous()))
        P[name] = m

    def _pack_u2_encoder(self, sd, P, pre, name):
        for i, scale in enumerate((4, 3, 2, 1)):
            self._pack_module(sd, P, f"{pre}.meta_unet_list.{i}", f"{name}{i}", scale, False, 2 if i == 0 else 1)
        P[f"{name}_last"] = (self._pack_gate(sd, f"{pre}.last_conv.0.conv", False, 1),
                             self._norm_params(sd, f"{pre}.last_conv.1"), sd[f"{pre}.last_conv.2.weight"].contiguous())

    def _pack(self):
        sd = self._sd()
        dev = next(iter(sd.values())).device
        P = {}
        self._pack_u2_encoder(sd, P, "zeroorderblock.en", "zen")
        self._pack_u2_encoder(sd, P, "separate_en", "sen")
        for i, scale in enumerate((1, 2, 3, 4)):
            self._pack_module(sd, P, f"zeroorderblock.de.meta_unet_list.{i}", f"zde{i}", scale, True, 1)
        pre = "zeroorderblock.de.last_conv"
        P["zde_last"] = (self._pack_gate(sd, pre + ".0.conv.0", True, 2), self._norm_params(sd, pre + ".1"),
                         sd[pre + ".2.weight"].contiguous(),
                         packing.pad_cols(sd[pre + ".3.weight"].reshape(1, 16).t().contiguous()), sd[pre + ".3.bias"].contiguous())
        for i in range(self.p):
            for j, d in enumerate(DILATIONS):
                self._pack_tcm(sd, P, f"zeroorderblock.tcms.{i}.tcm_list.{j}", f"ztcm{i}_{j}", d, ("left_conv", "right_conv"),
                               conv_idx=3, fir_idx=None)
        q = torch.arange(256, device=dev)
        ref_of_q = (q % 64) * 4 + q // 64              # channels-last feature f*64 + c  <-  reference feature c*4 + f
        for k in range(self.order_num):
            hp = f"highorderblock_list.{k}"
            w = sd[hp + ".in_conv.weight"][:, :, 0][ref_of_q]                                  # [256 (mine), 578]
            w_head = w[:, :256][:, ref_of_q].contiguous()
            w_ri = torch.zeros(256, RI_LD, device=dev)
            w_ri[:, :2 * N_BINS] = w[:, 256:]
            w_res = torch.zeros(RI_LD, 256, device=dev)
            w_res[:N_BINS] = sd[hp + ".real_resi.weight"][:, :, 0][:, ref_of_q]
            w_res[N_BINS:2 * N_BINS] = sd[hp + ".imag_resi.weight"][:, :, 0][:, ref_of_q]
            b_res = torch.zeros(RI_LD, device=dev)
            b_res[:N_BINS], b_res[N_BINS:2 * N_BINS] = sd[hp + ".real_resi.bias"], sd[hp + ".imag_resi.bias"]
            P[f"ho{k}"] = {"head": packing.split_tf32(w_head), "ri": packing.split_tf32(w_ri.contiguous()),
                           "bias": sd[hp + ".in_conv.bias"][ref_of_q].contiguous(),
                           "res": packing.split_tf32(w_res.contiguous()), "res_bias": b_res}
            for i in range(self.p):
                for j, d in enumerate(DILATIONS):
                    self._pack_tcm(sd, P, f"{hp}.tcms.{i}.tcm_list.{j}", f"htcm{k}_{i}_{j}", d, ("left_conv", "right_conv"),
                                   conv_idx=3, fir_idx=None)
        self._packed = P

    # -- runners ---------------------------------------------------------------------------------
    def _tcms(self, x, b, t, prefix):
        for i in range(self.p):
            for j in range(len(DILATIONS)):
                x = self._tcm(x, b, t, f"{prefix}{i}_{j}")
        return x

    @torch.no_grad()
    def forward(self, inputs, taps=None):
        if not inputs.is_cuda:
            raise RuntimeError("TaylorSENet (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(inputs, taps)

    def _forward_impl(self, inputs, taps=None):
        assert inputs.dim() == 4 and inputs.shape[1] == 2 and inputs.shape[3] == N_BINS, tuple(inputs.shape)
        b, _, t, _ = inputs.shape
        rows = self.forward_nhwc(inputs.float().permute(0, 2, 3, 1).contiguous(), taps).view(b, t, RI_LD)
        return torch.stack((rows[..., :N_BINS], rows[..., N_BINS:2 * N_BINS]), dim=1)

    def forward_nhwc(self, x, taps=None):
        """x [B,T,161,2] channels-last RI -> RI rows [B*T, 352] = [re(161) | im(161) | 0]."""
        self._ensure_packed()
        P = self._packed
        b, t = x.shape[0], x.shape[1]
        # ---- zeroth order (:139-153): U2 encoder -> TCMs -> U2 decoder -> gain ----
        en = self._u2_encoder(x, b, t, "zen", True)
        e5 = en[4]
        s = (e5.f32.view(b * t, 256), (e5.pair[0].view(b * t, 256), e5.pair[1].view(b * t, 256)))
        s = self._tcms(s, b, t, "ztcm")
        d, f = Act(None, (s[1][0].view(b, t, 4, 64), s[1][1].view(b, t, 4, 64))), 4
        for i in range(4):
            d, f = self._module(d, en[4 - i], b, t, f, f"zde{i}", 3)
        packed, (gamma, beta), slope, w1, b1 = P["zde_last"]
        tmp, f = self._gate_conv(d, en[0], b, t, f, packed, True, 5)
        g16 = self._gated_block(tmp, b, t, f, 16, gamma, beta, slope, want_f32=True, want_pair=False)
        gain = ops.linear(g16.f32.view(b * t * N_BINS, 16), w1, b1, 1, act="sigmoid").view(b, t, N_BINS)
        if taps is not None:
            taps["gain"] = gain
        zero, zero_pair = ops.taylor_zero(x, gain, RI_LD)                                     # :73-76
        # ---- high orders (:79-93) ----
        head = self._u2_encoder(x, b, t, "sen", False)[4]
        head_pair = (head.pair[0].view(b * t, 256), head.pair[1].view(b * t, 256))
        if taps is not None:
            taps["head"] = (head.pair[0] + head.pair[1]).view(b, t, 4, 64)
        out, pre, pre_pair = zero, zero, zero_pair
        for k in range(self.order_num):
            hp = P[f"ho{k}"]
            u1, _ = ops.gemm_tf32x3_ex(head_pair, hp["head"][0], hp["head"][1], None, 256)
            s = ops.gemm_tf32x3_ex(pre_pair, hp["ri"][0], hp["ri"][1], hp["bias"], 256, res=u1, want_f32=True,
                                   want_pair=True)                                            # in_conv over cat(head, pre)
            s = self._tcms(s, b, t, f"htcm{k}_")
            last = k == self.order_num - 1
            if k == 0:
                upd, upd_pair = ops.gemm_tf32x3_ex(s[1], hp["res"][0], hp["res"][1], hp["res_bias"], RI_LD,
                                                   want_f32=True, want_pair=not last)
            else:
                resi, _ = ops.gemm_tf32x3_ex(s[1], hp["res"][0], hp["res"][1], hp["res_bias"], RI_LD)
                upd, upd_pair = ops.axpby(resi, pre, 1.0, float(k), want_f32=True, want_pair=not last)   # + k * pre_term
            out, _ = ops.axpby(out, upd, 1.0, 1.0 / math.factorial(k + 1))
            pre, pre_pair = upd, upd_pair
        return out
