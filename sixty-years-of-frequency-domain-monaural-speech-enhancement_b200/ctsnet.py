"""Drop-in CTSNet ``Step1_net`` / ``Step2_net`` (reference: CTSNet/Step1_network.py:12-210,
CTSNet/Step2_network.py:13-209) and their causal ``_new`` twins
(CTSNet_new/*: InstanceNorm -> CumulativeLayerNorm, ``cumulative=True``).

    Step1_net().forward(mag [B,T,161])           -> magnitude estimate [B,T,161]
    Step2_net(X=6, R=3).forward(x [B,4,T,161])   -> complex residual   [B,2,T,161]

Same class names, constructors and state-dict keys as the reference, so the shipped
``BEST_MODEL/step{1,2}_*_cts_*_model*.pth`` load unchanged (two_stage_com_decode_vb.py:13-18).  Inference only.

How the reference's modules map onto the kernels (activations are channels-last [B,T,F,C]):
  * Gate_Conv (Step1_network.py:127-151), conv(x) * sigmoid(gate_conv(x)): the two convolutions are ONE implicit
    GEMM with outputs [a (C) | b (C)]; transposed convs are even / odd output-column parity classes with the
    Chomp_T(1) folded into the tap table (dt = -kt); skip ``torch.cat`` is a second source pointer.
  * InstanceNorm2d(affine) + PReLU(C) after every Gate_Conv: statistics per (clip, channel) over T x F of the
    gated product (se_chan_stats with the gate applied on the fly), then one normalise + PReLU pass
    (se_chan_norm) that also emits the TF32 split the next tensor-core layer reads.
  * TCM ``Glu`` / ``glu`` (Step1_network.py:163-193): in_conv and out_conv (+ residual) are tensor-core GEMMs over
    [B*T, 256]; the two branches (PReLU -> InstanceNorm1d -> ShareSepConv) read the same 64-channel tensor and
    are produced side by side as one 128-channel tensor (se_chan_norm, FIR post-op), the two dilated k=5 convs
    are ONE 5-tap implicit GEMM with a block-diagonal weight whose output [l | r] is exactly the gate layout
    of the output path (l * sigmoid(r) -> PReLU -> norm, applied on the fly by se_chan_stats / se_chan_norm).
  * the (c, f) flatten of the reference ([B, 64*4, T], :26-27) vs the channels-last flatten (f, c) is a
    one-time permutation of in_conv columns / out_conv rows.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import conv_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import build_param_tree

_F = [161, 79, 39, 19, 9, 4]
N_BINS = 161


def _norm_rows(pre, c, cumulative, dims):
    if cumulative:
        shape = (1, c) + (1,) * dims
        return [(pre + ".gain", shape, "param"), (pre + ".bias", shape, "param")]
    return [(pre + ".weight", (c,), "param"), (pre + ".bias", (c,), "param")]


def _codec_rows(en_pre, de_pres, cin, cumulative):
    rows = []
    for i in range(5):
        ci, kf = (cin, 5) if i == 0 else (64, 3)
        for n in ("conv", "gate_conv"):
            rows += [(f"{en_pre}.{i}.0.{n}.1.weight", (64, ci, 2, kf), "param"), (f"{en_pre}.{i}.0.{n}.1.bias", (64,), "param")]
        rows += _norm_rows(f"{en_pre}.{i}.1", 64, cumulative, 2)
        rows += [(f"{en_pre}.{i}.2.weight", (64,), "param")]
    for de_pre, fc, fc_first in de_pres:
        fc_rows = [(fc + ".weight", (161, 161), "param"), (fc + ".bias", (161,), "param")]
        if fc_first:
            rows += fc_rows
        for i in range(5):
            co, kf = (1, 5) if i == 4 else (64, 3)
            for n in ("conv", "gate_conv"):
                rows += [(f"{de_pre}.{i}.0.{n}.0.weight", (128, co, 2, kf), "param"),
                         (f"{de_pre}.{i}.0.{n}.0.bias", (co,), "param")]
            rows += _norm_rows(f"{de_pre}.{i}.1", co, cumulative, 2)
            rows += [(f"{de_pre}.{i}.2.weight", (co,), "param")]
        if not fc_first:
            rows += fc_rows
    return rows


def _tcm_rows(pre, j, branches, cumulative):
    rows = [(f"{pre}.in_conv.weight", (64, 256, 1), "param")]
    for n in branches:
        rows += [(f"{pre}.{n}.0.weight", (64,), "param")]
        rows += _norm_rows(f"{pre}.{n}.1", 64, cumulative, 1)
        rows += [(f"{pre}.{n}.2.weight", (1, 1, 2 * 2 ** j - 1), "param"), (f"{pre}.{n}.4.weight", (64, 64, 5), "param")]
    rows += [(f"{pre}.out_conv.0.weight", (64,), "param")]
    rows += _norm_rows(f"{pre}.out_conv.1", 64, cumulative, 1)
    rows += [(f"{pre}.out_conv.2.weight", (256, 64, 1), "param")]
    return rows


class _CtsBase(nn.Module):
    """Shared machinery of the two stages: packing, encoder / TCM / decoder runners."""

    def __init__(self, cumulative):
        super().__init__()
        self.cumulative = bool(cumulative)
        self._packed = None
        self._packed_key = None

    # -- weight packing --------------------------------------------------------------------------
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

    def _sd(self):
        return {k: v.detach().float() for k, v in self.state_dict().items()}

    def _norm_params(self, sd, pre):
        g = sd[pre + (".gain" if self.cumulative else ".weight")].reshape(-1).contiguous()
        return g, sd[pre + ".bias"].reshape(-1).contiguous()

    def _pack_encoder(self, sd, P, en_pre):
        for i in range(5):
            w1, w2 = sd[f"{en_pre}.{i}.0.conv.1.weight"], sd[f"{en_pre}.{i}.0.gate_conv.1.weight"]   # [64, Ci, 2, kf]
            kf = w1.shape[-1]
            w = torch.cat([torch.cat([w1[:, :, kt, k].t(), w2[:, :, kt, k].t()], 1)
                           for kt in range(2) for k in range(kf)], 0)                        # K = (kt, kf, ci), N = [a | b]
            bias = torch.cat([sd[f"{en_pre}.{i}.0.conv.1.bias"], sd[f"{en_pre}.{i}.0.gate_conv.1.bias"]]).contiguous()
            P[f"enc{i}"] = (ConvWeights(w.contiguous(), 128), bias, [(kt - 1, k) for kt in range(2) for k in range(kf)],
                            *self._norm_params(sd, f"{en_pre}.{i}.1"), sd[f"{en_pre}.{i}.2.weight"].contiguous())

    def _pack_decoder(self, sd, P, de_pre, fc, name):
        for i in range(5):
            w1, w2 = sd[f"{de_pre}.{i}.0.conv.0.weight"], sd[f"{de_pre}.{i}.0.gate_conv.0.weight"]   # [128, Co, 2, kf]
            co, kf = w1.shape[1], w1.shape[-1]
            tap = lambda kt, k: torch.cat([w1[:, :, kt, k], w2[:, :, kt, k]], 1)             # noqa: E731  [128, 2Co]
            ev = [(kt, k) for kt in range(2) for k in range(0, kf, 2)]
            od = [(kt, k) for kt in range(2) for k in range(1, kf, 2)]
            # out[t, f'] += in[t - kt, (f' - k) / 2] W[kt, k]:   f' = 2m + (k & 1)  ->  f = m - k // 2
            even = ConvWeights(torch.cat([tap(kt, k) for kt, k in ev], 0).contiguous(), 2 * co)
            odd = ConvWeights(torch.cat([tap(kt, k) for kt, k in od], 0).contiguous(), 2 * co)
            bias = torch.cat([sd[f"{de_pre}.{i}.0.conv.0.bias"], sd[f"{de_pre}.{i}.0.gate_conv.0.bias"]]).contiguous()
            P[f"{name}{i}"] = (even, odd, bias, [(-kt, -(k // 2)) for kt, k in ev], [(-kt, -(k // 2)) for kt, k in od],
                               *self._norm_params(sd, f"{de_pre}.{i}.1"), sd[f"{de_pre}.{i}.2.weight"].contiguous())
        P[f"{name}_fc"] = (packing.pad_cols(sd[fc + ".weight"].t().contiguous()), sd[fc + ".bias"].contiguous())

    def _pack_tcm(self, sd, P, pre, name, d, branches, conv_idx=4, fir_idx=2):
        """conv_idx / fir_idx: positions of the dilated conv and of ShareSepConv inside the branch nn.Sequential
        (CTSNet: 4 / 2; TaylorSENet's SqueezedTCM has no ShareSepConv: 3 / None)."""
        dev = sd[f"{pre}.in_conv.weight"].device
        q = torch.arange(256, device=dev)
        ref_of_q = (q % 64) * 4 + q // 64              # channels-last column f*64 + c  <-  reference feature c*4 + f
        w_in = sd[f"{pre}.in_conv.weight"][:, :, 0][:, ref_of_q].contiguous()                # [64, 256]  (N, K)
        w_out = sd[f"{pre}.out_conv.2.weight"][:, :, 0][ref_of_q].contiguous()              # [256, 64]
        wl, wr = sd[f"{pre}.{branches[0]}.{conv_idx}.weight"], sd[f"{pre}.{branches[1]}.{conv_idx}.weight"]    # [64, 64, 5]
        wd = torch.zeros(128, 5, 128, device=dev)                                            # [co, tap, ci] block diagonal
        wd[:64, :, :64] = wl.permute(0, 2, 1)
        wd[64:, :, 64:] = wr.permute(0, 2, 1)
        gl, bl = self._norm_params(sd, f"{pre}.{branches[0]}.1")
        gr, br = self._norm_params(sd, f"{pre}.{branches[1]}.1")
        go, bo = self._norm_params(sd, f"{pre}.out_conv.1")
        fir = None
        if fir_idx is not None:
            fir = torch.stack([sd[f"{pre}.{branches[0]}.{fir_idx}.weight"].reshape(-1),
                               sd[f"{pre}.{branches[1]}.{fir_idx}.weight"].reshape(-1)]).contiguous()
        P[name] = {
            "in": packing.split_tf32(w_in), "out": packing.split_tf32(w_out),
            "dil": packing.split_tf32(wd.reshape(128, 640).contiguous()),
            "taps": [((j - 4) * d, 0) for j in range(5)],
            "slope_lr": torch.cat([sd[f"{pre}.{branches[0]}.0.weight"], sd[f"{pre}.{branches[1]}.0.weight"]]).contiguous(),
            "gamma_lr": torch.cat([gl, gr]).contiguous(), "beta_lr": torch.cat([bl, br]).contiguous(),
            "fir": fir,
            "slope_o": sd[f"{pre}.out_conv.0.weight"].contiguous(), "gamma_o": go, "beta_o": bo,
        }

    # -- runners ---------------------------------------------------------------------------------
    def _stats(self, x, b, t, f, c, pre, slope=None, groups=1):
        if self.cumulative:
            return ops.cum_stats(x, b, t, f, c, pre, slope, groups=groups)
        return ops.chan_stats(x, b, t * f, c, pre, slope)

    def _norm(self, x, b, t, f, c, st, gamma, beta, pre, slope=None, groups=1, **kw):
        return ops.chan_norm(x, b, t * f, c, st[0], st[1], gamma, beta, pre=pre, pre_slope=slope,
                             cumulative=self.cumulative, rows_per_t=f, stat_groups=groups, **kw)

    def _gated_block(self, tmp, b, t, fout, co, gamma, beta, slope, want_f32, want_pair):
        """InstanceNorm / cLN + PReLU of the gated product held in tmp [B,T,fout,2*co]."""
        st = self._stats(tmp, b, t, fout, co, "glu")
        f32, pair = self._norm(tmp, b, t, fout, co, st, gamma, beta, "glu", post="prelu", post_slope=slope,
                               want_f32=want_f32, want_pair=want_pair)
        v = lambda z: z.view(b, t, fout, co)      # noqa: E731
        return Act(v(f32) if f32 is not None else None, (v(pair[0]), v(pair[1])) if pair is not None else None)

    def _encoder(self, x, taps):
        """x [B,T,161,Cin] fp32 -> list of 5 Acts."""
        P = self._packed
        b, t = x.shape[0], x.shape[1]
        h = Act(x)
        outs = []
        for i in range(5):
            w, bias, tp, gamma, beta, slope = P[f"enc{i}"]
            fin, fout = _F[i], _F[i + 1]
            tmp = Act(torch.empty(b, t, fout, 128, device=x.device, dtype=torch.float32))
            conv_engine.conv(h, None, b, t, fin, fout, tp, 2, w, bias, "none", tmp, fout)
            # e1 also feeds the last decoder layer (fp32 FMA engine), e5 is the fp32 residual of the first TCM
            h = self._gated_block(tmp.f32, b, t, fout, 64, gamma, beta, slope, want_f32=(i in (0, 4)), want_pair=True)
            outs.append(h)
            if taps is not None:
                taps[f"e{i + 1}"] = h.pair[0] + h.pair[1]
        return outs

    def _tcm(self, x, b, t, name):
        """One Glu block on the residual stream x = (f32 [B*T,256], pair)."""
        p = self._packed[name]
        xf, xp = x
        u, _ = ops.gemm_tf32x3_ex(xp, p["in"][0], p["in"][1], None, 64)                        # in_conv
        st = self._stats(u, b, t, 1, 128, "prelu", p["slope_lr"], groups=2)
        post = dict(post="fir", fir_w=p["fir"], fir_groups=2) if p["fir"] is not None else dict(post="none")
        _, y = self._norm(u, b, t, 1, 128, st, p["gamma_lr"], p["beta_lr"], "prelu", p["slope_lr"], groups=2,
                          want_f32=False, want_pair=True, **post)
        v = torch.empty(b, t, 1, 128, device=u.device, dtype=torch.float32)                  # [l | r]
        ops.conv_tf32x3((y[0].view(b, t, 1, 128), y[1].view(b, t, 1, 128)), None, b, t, 1, 1, p["taps"], 1,
                        p["dil"][0], p["dil"][1], None, 128, "none", 1, out=v)
        st = self._stats(v, b, t, 1, 64, "glu_prelu", p["slope_o"])
        _, z = self._norm(v, b, t, 1, 64, st, p["gamma_o"], p["beta_o"], "glu_prelu", p["slope_o"], want_f32=False,
                          want_pair=True)
        return ops.gemm_tf32x3_ex((z[0].view(b * t, 64), z[1].view(b * t, 64)), p["out"][0], p["out"][1], None, 256,
                                  res=xf, want_f32=True, want_pair=True)                      # out_conv + residual

    def _tcm_stack(self, e5, b, t, names, taps):
        """names: one list of TCM blocks per repeat; returns the sum of the repeats' outputs (x_acc,
        Step1_network.py:28-33) as an Act [B,T,4,64] (TF32 pair for the decoder's tensor-core convs)."""
        x = (e5.f32.view(b * t, 256), (e5.pair[0].view(b * t, 256), e5.pair[1].view(b * t, 256)))
        acc = pr = None
        for ri, group in enumerate(names):
            for nm in group:
                x = self._tcm(x, b, t, nm)
            last = ri == len(names) - 1
            if acc is None:
                acc, pr = x
            else:
                acc, pr = ops.add(acc, x[0], want_f32=(not last) or taps is not None, want_pair=last)
        if taps is not None:
            taps["tcm"] = acc.view(b, t, 4, 64)
        return Act(None, (pr[0].view(b, t, 4, 64), pr[1].view(b, t, 4, 64)))

    def _decoder(self, d, enc, b, t, name, act):
        P = self._packed
        for i in range(5):
            even, odd, bias, tp_e, tp_o, gamma, beta, slope = P[f"{name}{i}"]
            fin = _F[5 - i]
            fout = _F[4 - i]
            co = even.cout // 2
            ne, no = (fout + 1) // 2, fout // 2
            skip = enc[4 - i]
            tmp = Act(torch.empty(b, t, fout, 2 * co, device=skip.pair[0].device, dtype=torch.float32))
            conv_engine.conv(d, skip, b, t, fin, ne, tp_e, 1, even, bias, "none", tmp, fout, dst_f0=0, dst_fstep=2)
            conv_engine.conv(d, skip, b, t, fin, no, tp_o, 1, odd, bias, "none", tmp, fout, dst_f0=1, dst_fstep=2)
            # layer 4 (Cout = 1 + gate) runs on the fp32 FMA engine: its inputs (layer 3 output, e1) stay fp32
            d = self._gated_block(tmp.f32, b, t, fout, co, gamma, beta, slope, want_f32=(i >= 3), want_pair=(i < 3))
        wfc, bfc = P[f"{name}_fc"]
        return ops.linear(d.f32.view(b * t, N_BINS), wfc, bfc, N_BINS, act=act).view(b, t, N_BINS)


class Step1_net(_CtsBase):
    def __init__(self, cumulative=False):
        super().__init__(cumulative)
        rows = _codec_rows("en.en", [("de.de", "de.de6.0", True)], 1, cumulative)
        for s in (1, 2, 3):
            for j in range(6):
                rows += _tcm_rows(f"tcm{s}.tcm_list.{j}", j, ("left_conv", "right_conv"), cumulative)
        build_param_tree(self, rows)

    def _pack(self):
        sd = self._sd()
        P = {}
        self._pack_encoder(sd, P, "en.en")
        self._pack_decoder(sd, P, "de.de", "de.de6.0", "dec")
        for s in (1, 2, 3):
            for j in range(6):
                self._pack_tcm(sd, P, f"tcm{s}.tcm_list.{j}", f"tcm{s}_{j}", 2 ** j, ("left_conv", "right_conv"))
        self._packed = P

    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("CTSNet Step1_net (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        assert x.dim() == 3 and x.shape[2] == N_BINS, tuple(x.shape)
        self._ensure_packed()
        b, t = x.shape[0], x.shape[1]
        enc = self._encoder(x.float().contiguous().view(b, t, N_BINS, 1), taps)
        d = self._tcm_stack(enc[4], b, t, [[f"tcm{s}_{j}" for j in range(6)] for s in (1, 2, 3)], taps)
        return self._decoder(d, enc, b, t, "dec", "softplus")


class Step2_net(_CtsBase):
    def __init__(self, X=6, R=3, cumulative=False):
        super().__init__(cumulative)
        self.X, self.R = X, R
        rows = _codec_rows("en.en_module", [("de_r.de_list", "de_r.de6.0", False), ("de_i.de_list", "de_i.de6.0", False)],
                           4, cumulative)
        for r in range(R):
            for j in range(X):
                rows += _tcm_rows(f"tcm_list.{r}.glu_list.{j}", j, ("ori_conv", "att_ori"), cumulative)
        build_param_tree(self, rows)

    def _pack(self):
        sd = self._sd()
        P = {}
        self._pack_encoder(sd, P, "en.en_module")
        self._pack_decoder(sd, P, "de_r.de_list", "de_r.de6.0", "dec_r")
        self._pack_decoder(sd, P, "de_i.de_list", "de_i.de6.0", "dec_i")
        for r in range(self.R):
            for j in range(self.X):
                self._pack_tcm(sd, P, f"tcm_list.{r}.glu_list.{j}", f"tcm{r}_{j}", 2 ** j, ("ori_conv", "att_ori"))
        self._packed = P

    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("CTSNet Step2_net (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        assert x.dim() == 4 and x.shape[1] == 4 and x.shape[3] == N_BINS, tuple(x.shape)
        yr, yi = self.forward_nhwc(x.float().permute(0, 2, 3, 1).contiguous(), taps)
        return torch.stack((yr, yi), dim=1)

    def forward_nhwc(self, x, taps=None):
        """x [B,T,161,4] channels-last (noisy re, im, stage-1 re, im) -> (real [B,T,161], imag [B,T,161])."""
        self._ensure_packed()
        b, t = x.shape[0], x.shape[1]
        enc = self._encoder(x, taps)
        d = self._tcm_stack(enc[4], b, t, [[f"tcm{r}_{j}" for j in range(self.X)] for r in range(self.R)], taps)
        return self._decoder(d, enc, b, t, "dec_r", "none"), self._decoder(d, enc, b, t, "dec_i", "none")
