"""Drop-in G2Net ``gaf_base`` (reference: G2Net_new/gaf_net_320.py:10-87 -- cumulative LayerNorm -- and
G2Net_VB/gaf_net_320.py -- InstanceNorm).

    gaf_base(3, 64, 2, 4, 4, [1, 2, 5, 9], 256 + 161 * 2, 256, 256, (2, 3), (1, 3), 64, 'cat', 3, is_aux=False,
             encoder_type='U2Net', tcm_type='full-band').forward(x [B,2,T,161]) -> list of 3 x [B,2,161,T]

i.e. the configuration com_decode.py:23 builds; the 825 state-dict keys are the reference's, so the shipped
``BEST_MODEL/vb_gaf_*_model.pth`` load unchanged.  ``cumulative=False`` selects the G2Net_VB norms.  Inference only.

Mapping onto the kernels (shared machinery: ctsnet.py / taylor.py):
  * U2Net_Encoder (:277-303): En_unet_module = gated (2,3) conv pair as ONE implicit GEMM [a | b] + norm + PReLU, inner
    U-Net of k(1,3) Conv2dunit / Deconv2dunit levels, residual.
  * GAF_module (:90-115): cat(feat_x, flatten(pre_x)) never exists -- every 578-wide 1x1 conv is two tensor-core GEMMs
    (K = 256 over the encoder feature, K = 352 over the "RI row" of pre_x, the second adding the first in its
    epilogue); main and gate convs of a branch are one GEMM with N = 512 followed by the gate pass; the glance and
    focus heads write gain / residual, and se_gaf_update forms gain * |pre| * e^{j angle pre} + residual as the next
    RI row (fp32 + TF32 pair).
  * Glu (:245-274), single-branch squeezed TCM: GEMM 256->64, PReLU + norm (two-pass), 3-tap dilated conv as implicit
    GEMM, PReLU + norm, GEMM 64->256 + residual.
"""
from __future__ import annotations

import json
import os

import torch

from . import ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import build_param_tree
from .taylor import _U2Base

N_BINS = 161
RI_LD, IM_OFF = 352, 176          # RI row: re at 0, im at 176 (16-byte aligned GEMM outputs), zero padded to 11 x 32
DILAS = (1, 2, 5, 9)
_HERE = os.path.dirname(os.path.abspath(__file__))

ENC13 = [(0, 0), (0, 1), (0, 2)]
DEC13_EVEN = [(0, 0), (0, -1)]
DEC13_ODD = [(0, 0)]


def _keys(cumulative):
    with open(os.path.join(_HERE, "g2net_new_keys.json" if cumulative else "g2net_vb_keys.json")) as f:
        return json.load(f)


class gaf_base(_U2Base):
    def __init__(self, kd1=3, cd1=64, tcm_num=2, sub_g1=4, sub_g2=4, dilas=(1, 2, 5, 9), ci=256 + 161 * 2, co1=256, co2=256,
                 k1=(2, 3), k2=(1, 3), c=64, intra_connect="cat", stage_num=3, is_causal=True, is_aux=True,
                 encoder_type="U2Net", tcm_type="full-band", cumulative=True):
        super().__init__(cumulative)
        cfg = (kd1, cd1, tcm_num, tuple(dilas), ci, co1, co2, tuple(k1), tuple(k2), c, intra_connect, stage_num, is_causal,
               is_aux, encoder_type, tcm_type)
        if cfg != (3, 64, 2, DILAS, 578, 256, 256, (2, 3), (1, 3), 64, "cat", 3, True, False, "U2Net", "full-band"):
            raise NotImplementedError("se_b200.g2net.gaf_base implements the configuration of G2Net_new/com_decode.py:23")
        self.stage_num, self.tcm_num = stage_num, tcm_num
        build_param_tree(self, [(k, tuple(s), "param") for k, s in _keys(self.cumulative).items()])

    # -- weight packing --------------------------------------------------------------------------
    def _pack_gate_pair(self, sd, pre):
        """Gate_2dconv (:465-486), de_flag False: conv / gate_conv with a causal top pad -> one [a | b] GEMM."""
        w1, w2 = sd[f"{pre}.conv.1.weight"], sd[f"{pre}.gate_conv.1.weight"]                 # [64, Ci, 2, kf]
        kf = w1.shape[-1]
        w = torch.cat([torch.cat([w1[:, :, kt, k].t(), w2[:, :, kt, k].t()], 1) for kt in range(2) for k in range(kf)], 0)
        bias = torch.cat([sd[f"{pre}.conv.1.bias"], sd[f"{pre}.gate_conv.1.bias"]]).contiguous()
        return [(ConvWeights(w.contiguous(), 128), [(kt - 1, k) for kt in range(2) for k in range(kf)])], bias

    def _pack_g2_module(self, sd, P, pre, name, scale):
        m = {"in": self._pack_gate_pair(sd, f"{pre}.in_conv.0"), "in_norm": self._norm_params(sd, f"{pre}.in_conv.1"),
             "in_slope": sd[f"{pre}.in_conv.2.weight"].contiguous(), "scale": scale, "transpose": False,
             "enco": [], "deco": []}
        for j in range(scale):
            w = sd[f"{pre}.enco.{j}.conv.0.weight"]                                            # [64, 64, 1, 3]
            wk = w.permute(2, 3, 1, 0).reshape(-1, 64).contiguous()
            m["enco"].append((ConvWeights(wk, 64), sd[f"{pre}.enco.{j}.conv.0.bias"].contiguous(),
                              *self._norm_params(sd, f"{pre}.enco.{j}.conv.1"), sd[f"{pre}.enco.{j}.conv.2.weight"].contiguous()))
            w = sd[f"{pre}.deco.{j}.deconv.0.weight"]                                          # [64 | 128, 64, 1, 3]
            even = torch.cat([w[:, :, 0, 0], w[:, :, 0, 2]], 0).contiguous()
            odd = w[:, :, 0, 1].contiguous()
            m["deco"].append((ConvWeights(even, 64), ConvWeights(odd, 64), sd[f"{pre}.deco.{j}.deconv.0.bias"].contiguous(),
                              *self._norm_params(sd, f"{pre}.deco.{j}.deconv.1"),
                              sd[f"{pre}.deco.{j}.deconv.2.weight"].contiguous()))
        P[name] = m

    def _pack_glu(self, sd, P, pre, name, d, ref_of_q):
        w_in = sd[f"{pre}.in_conv.weight"][:, :, 0][:, ref_of_q].contiguous()                 # [64, 256]
        w_out = sd[f"{pre}.out_conv.2.weight"][:, :, 0][ref_of_q].contiguous()               # [256, 64]
        wd = sd[f"{pre}.left_conv.3.weight"].permute(0, 2, 1).reshape(64, 192).contiguous()   # [co, tap*64 + ci]
        P[name] = {"in": packing.split_tf32(w_in), "out": packing.split_tf32(w_out), "dil": packing.split_tf32(wd),
                   "taps": [((j - 2) * d, 0) for j in range(3)],
                   "slope_l": sd[f"{pre}.left_conv.0.weight"].contiguous(), "norm_l": self._norm_params(sd, f"{pre}.left_conv.1"),
                   "slope_o": sd[f"{pre}.out_conv.0.weight"].contiguous(), "norm_o": self._norm_params(sd, f"{pre}.out_conv.1")}

    def _pack(self):
        sd = self._sd()
        dev = next(iter(sd.values())).device
        P = {}
        for i, scale in enumerate((4, 3, 2, 1)):
            self._pack_g2_module(sd, P, f"en.meta_unet_list.{i}", f"en{i}", scale)
        P["en_last"] = (self._pack_gate_pair(sd, "en.last_conv.0"), self._norm_params(sd, "en.last_conv.1"),
                        sd["en.last_conv.2.weight"].contiguous())
        q = torch.arange(256, device=dev)
        ref_of_q = (q % 64) * 4 + q // 64              # channels-last feature f*64 + c  <-  reference feature c*4 + f

        def ri_cols(w):                                 # [N, 322] over (ri*161 + f)  ->  [N, 352] over the RI-row layout
            out = torch.zeros(w.shape[0], RI_LD, device=dev)
            out[:, :N_BINS] = w[:, :N_BINS]
            out[:, IM_OFF:IM_OFF + N_BINS] = w[:, N_BINS:]
            return out.contiguous()

        for s in range(self.stage_num):
            for br, heads in (("glance_branch", ("mstcm_filter",)), ("focus_branch", ("mstcm_r", "mstcm_i"))):
                pre = f"gafs.{s}.{br}"
                # main / gate in_convs: rows in MY feature order, stacked [main (256) | gate (256)]
                w = torch.cat([sd[pre + ".in_conv_main.weight"][:, :, 0][ref_of_q],
                               sd[pre + ".in_conv_gate.0.weight"][:, :, 0][ref_of_q]], 0)        # [512, 578]
                bias = torch.cat([sd[pre + ".in_conv_main.bias"][ref_of_q], sd[pre + ".in_conv_gate.0.bias"][ref_of_q]])
                P[f"g{s}_{br}_in"] = {"feat": packing.split_tf32(w[:, :256][:, ref_of_q].contiguous()),
                                      "ri": packing.split_tf32(ri_cols(w[:, 256:])), "bias": bias.contiguous()}
                for h in heads:
                    for i in range(self.tcm_num):
                        for j, d in enumerate(DILAS):
                            self._pack_glu(sd, P, f"{pre}.{h}.{i}.tcm_list.{j}", f"g{s}_{h}_{i}_{j}", d, ref_of_q)
                    wf = sd[f"{pre}.{h}.{self.tcm_num}.weight"][:, :, 0][:, ref_of_q].contiguous()   # [161, 256]
                    P[f"g{s}_{h}_out"] = (packing.split_tf32(wf), sd[f"{pre}.{h}.{self.tcm_num}.bias"].contiguous())
        self._packed = P

    # -- runners ---------------------------------------------------------------------------------
    def _inner_taps(self):
        return ENC13, DEC13_EVEN, DEC13_ODD

    def _glu(self, x, b, t, name):
        """Glu (:268-274) on the residual stream x = (f32 [B*T,256], pair)."""
        p = self._packed[name]
        xf, xp = x
        u, _ = ops.gemm_tf32x3_ex(xp, p["in"][0], p["in"][1], None, 64)
        st = self._stats(u, b, t, 1, 64, "prelu", p["slope_l"])
        _, y = self._norm(u, b, t, 1, 64, st, *p["norm_l"], "prelu", p["slope_l"], want_f32=False, want_pair=True)
        v = torch.empty(b, t, 1, 64, device=u.device, dtype=torch.float32)
        ops.conv_tf32x3((y[0].view(b, t, 1, 64), y[1].view(b, t, 1, 64)), None, b, t, 1, 1, p["taps"], 1, p["dil"][0],
                        p["dil"][1], None, 64, "none", 1, out=v)
        st = self._stats(v, b, t, 1, 64, "prelu", p["slope_o"])
        _, z = self._norm(v, b, t, 1, 64, st, *p["norm_o"], "prelu", p["slope_o"], want_f32=False, want_pair=True)
        return ops.gemm_tf32x3_ex((z[0].view(b * t, 64), z[1].view(b * t, 64)), p["out"][0], p["out"][1], None, 256,
                                  res=xf, want_f32=True, want_pair=True)

    def _branch_in(self, feat_pair, pre_pair, name):
        """in_conv_main(x) * sigmoid(in_conv_gate(x)) over x = cat(feat, flatten(pre)) -> (f32, pair) [B*T, 256]."""
        p = self._packed[name]
        u1, _ = ops.gemm_tf32x3_ex(feat_pair, p["feat"][0], p["feat"][1], None, 512)
        u, _ = ops.gemm_tf32x3_ex(pre_pair, p["ri"][0], p["ri"][1], p["bias"], 512, res=u1)
        return ops.glu_affine_act(u, None, None, "none", want_f32=True, want_pair=True)

    def _head(self, x, b, t, s, h, out, act):
        for i in range(self.tcm_num):
            for j in range(len(DILAS)):
                x = self._glu(x, b, t, f"g{s}_{h}_{i}_{j}")
        (w_hi, w_lo), bias = self._packed[f"g{s}_{h}_out"]
        ops.gemm_tf32x3(x[1][0], x[1][1], w_hi, w_lo, bias, N_BINS, act=act, out=out)

    @torch.no_grad()
    def forward(self, inpt, taps=None):
        if not inpt.is_cuda:
            raise RuntimeError("G2Net gaf_base (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(inpt, taps)

    def _forward_impl(self, inpt, taps=None):
        assert inpt.dim() == 4 and inpt.shape[1] == 2 and inpt.shape[3] == N_BINS, tuple(inpt.shape)
        b, _, t, _ = inpt.shape
        outs = []
        for rows in self.forward_nhwc(inpt.float().permute(0, 2, 3, 1).contiguous(), taps, all_stages=True):
            r = rows.view(b, t, RI_LD)
            outs.append(torch.stack((r[..., :N_BINS], r[..., IM_OFF:IM_OFF + N_BINS]), dim=1).transpose(-2, -1).contiguous())
        return outs

    def forward_nhwc(self, x, taps=None, all_stages=False):
        """x [B,T,161,2] channels-last RI -> RI rows [B*T, 352] of the last stage (or the list of all stages)."""
        self._ensure_packed()
        b, t = x.shape[0], x.shape[1]
        rows = b * t
        feat = self._u2_encoder(x, b, t, "en", False)[4]
        feat_pair = (feat.pair[0].view(rows, 256), feat.pair[1].view(rows, 256))
        if taps is not None:
            taps["feat"] = (feat.pair[0] + feat.pair[1]).view(b, t, 4, 64)
        # pre_x of the first stage is the network input itself (:77)
        pre, pre_pair = ops.gaf_update(x, x[..., 1], 2 * N_BINS, 2, None, None, rows, N_BINS, RI_LD, IM_OFF)
        outs = []
        for s in range(self.stage_num):
            gain = torch.empty(rows, N_BINS, device=x.device, dtype=torch.float32)
            resi = torch.zeros(rows, RI_LD, device=x.device, dtype=torch.float32)
            xg = self._branch_in(feat_pair, pre_pair, f"g{s}_glance_branch_in")
            self._head(xg, b, t, s, "mstcm_filter", gain, "sigmoid")
            xf = self._branch_in(feat_pair, pre_pair, f"g{s}_focus_branch_in")
            self._head(xf, b, t, s, "mstcm_r", resi[:, :N_BINS], "none")
            self._head(xf, b, t, s, "mstcm_i", resi[:, IM_OFF:IM_OFF + N_BINS], "none")
            pre, pre_pair = ops.gaf_update(pre, pre[:, IM_OFF:], RI_LD, 1, gain, resi, rows, N_BINS, RI_LD, IM_OFF,
                                           want_pair=s < self.stage_num - 1)
            outs.append(pre)
        return outs if all_stages else outs[-1]
