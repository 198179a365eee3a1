"""Drop-in ``dpcrn`` (reference: DPCRN/DPCRN.py:16-186).

forward(x [B,2,T,161] real/imag planes) -> [B,2,T,161] (the complex ratio mask is applied inside forward,
DPCRN.py:33-42); same class name, constructor and state-dict keys as the reference, so
``dpcrn().load_state_dict(torch.load('BEST_MODEL/vb_dpcrn_noncprs_model.pth'))`` works unchanged
(DPCRN/dpcrn_decode_vb.py:18-22).  Inference only.

How the reference's modules map onto the kernels (activations are channels-last [B,T,F,C]):
  * encoder / decoder: the CRN conv engine (causal k(2,3) s(1,2) implicit GEMMs, eval BatchNorm folded, PReLU with
    its scalar slope in the GEMM epilogue, transposed convs as even/odd column parity classes, skip concat as
    two source pointers, de4's extra left pad as a BN(0) fill column; DPCRN.py:100-179).
  * DPRNN (:44-98), applied twice with the same weights (:28-29):
      intra Bi-LSTM over F = 4 (:70): 25.7 K sequences of length 4 -- the opposite regime of the persistent
        recurrence kernel.  Every (layer, direction, position) is ONE fused tensor-core LSTM-cell GEMM
        (se_lstm_cell_tf32x3_ex: [x_f | h] W^T, gates + state update in the epilogue) over all B*T rows, writing
        h straight into the [B*T, 4, 128] (fwd | bwd) buffer that is both the next step's state and the next
        layer's input; 16 launches per DPRNN pass.
      inter LSTM over T (:82): the F = 4 frequency positions are the GROUPS of one multi-group persistent
        recurrence launch with shared weights (se_lstm_seq_multi, whh_group_stride = 0); the (b,f) <-> (b,t)
        permutes of the reference (:80,:87) are just the group offsets of that launch.
      Linear + LayerNorm([4,128]) + residual (:72-76, :84-90): tensor-core GEMM, then se_group_layernorm over the
        512 (f,c) values of a frame with the residual fused.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import conv_engine, lstm_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import bn_rows, build_param_tree, lstm_rows

_ENC_CH = [2, 32, 32, 32, 64, 128]           # DPCRN.py:104-128
_ENC_F = [161, 80, 39, 19, 9, 4]
_DEC_CH = [(256, 64), (128, 32), (64, 32), (64, 32), (64, 2)]   # DPCRN.py:144-168


def _spec():
    rows = []
    for i in range(5):
        ci, co = _ENC_CH[i], _ENC_CH[i + 1]
        rows += [(f"en.en_module.{i}.1.weight", (co, ci, 2, 3), "param"), (f"en.en_module.{i}.1.bias", (co,), "param")]
        rows += bn_rows(f"en.en_module.{i}.2", co)
        rows += [(f"en.en_module.{i}.3.weight", (1,), "param")]
    for l in range(2):
        for sfx in ("", "_reverse"):
            rows += [(f"dprnn.intra_rnn.weight_ih_l{l}{sfx}", (256, 128), "param"),
                     (f"dprnn.intra_rnn.weight_hh_l{l}{sfx}", (256, 64), "param"),
                     (f"dprnn.intra_rnn.bias_ih_l{l}{sfx}", (256,), "param"),
                     (f"dprnn.intra_rnn.bias_hh_l{l}{sfx}", (256,), "param")]
    rows += [("dprnn.intra_fc.weight", (128, 128), "param"), ("dprnn.intra_fc.bias", (128,), "param")]
    rows += lstm_rows("dprnn.inter_rnn", 128, 128, 2)
    rows += [("dprnn.inter_fc.weight", (128, 128), "param"), ("dprnn.inter_fc.bias", (128,), "param")]
    rows += [("dprnn.ln1.weight", (4, 128), "param"), ("dprnn.ln1.bias", (4, 128), "param"),
             ("dprnn.ln2.weight", (4, 128), "param"), ("dprnn.ln2.bias", (4, 128), "param")]
    for i, (ci, co) in enumerate(_DEC_CH):
        rows += [(f"de.de_module.{i}.0.weight", (ci, co, 2, 3), "param"), (f"de.de_module.{i}.0.bias", (co,), "param")]
        if i < 4:
            bn = 3 if i == 3 else 2            # de4 has the extra pad module (DPCRN.py:159-165)
            rows += bn_rows(f"de.de_module.{i}.{bn}", co)
            rows += [(f"de.de_module.{i}.{bn + 1}.weight", (1,), "param")]
    return rows


class dpcrn(nn.Module):
    N_BINS = 161

    def __init__(self):
        super().__init__()
        build_param_tree(self, _spec())
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

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items() if v.is_floating_point()}
        P = {}
        for i in range(5):
            bn = tuple(sd[f"en.en_module.{i}.2.{n}"] for n in ("weight", "bias", "running_mean", "running_var"))
            w, bias = packing.pack_conv(sd[f"en.en_module.{i}.1.weight"], sd[f"en.en_module.{i}.1.bias"], bn)
            P[f"en{i}"] = (ConvWeights(w, _ENC_CH[i + 1]), bias, float(sd[f"en.en_module.{i}.3.weight"].item()))
        for l in range(2):
            for d, sfx in enumerate(("", "_reverse")):
                pre = "dprnn.intra_rnn."
                c = packing.pack_lstm_cell(sd[f"{pre}weight_ih_l{l}{sfx}"], sd[f"{pre}weight_hh_l{l}{sfx}"],
                                           sd[f"{pre}bias_ih_l{l}{sfx}"], sd[f"{pre}bias_hh_l{l}{sfx}"])
                P[f"intra{l}{d}"] = (c["w_hi"], c["w_lo"], c["bias"])
        for nm in ("intra_fc", "inter_fc"):
            w = sd[f"dprnn.{nm}.weight"]
            hi, lo = packing.split_tf32(w.contiguous())
            P[nm] = (hi, lo, sd[f"dprnn.{nm}.bias"].contiguous(), packing.pad_cols(w.t().contiguous()))
        for l in range(2):
            P[f"inter{l}"] = packing.pack_lstm_layer(sd[f"dprnn.inter_rnn.weight_ih_l{l}"], sd[f"dprnn.inter_rnn.weight_hh_l{l}"],
                                                     sd[f"dprnn.inter_rnn.bias_ih_l{l}"], sd[f"dprnn.inter_rnn.bias_hh_l{l}"])
        for nm in ("ln1", "ln2"):
            P[nm] = (sd[f"dprnn.{nm}.weight"].reshape(-1).contiguous(), sd[f"dprnn.{nm}.bias"].reshape(-1).contiguous())
        for i in range(5):
            w, b = sd[f"de.de_module.{i}.0.weight"], sd[f"de.de_module.{i}.0.bias"]
            if i < 4:
                bnm = 3 if i == 3 else 2
                bn = tuple(sd[f"de.de_module.{i}.{bnm}.{n}"] for n in ("weight", "bias", "running_mean", "running_var"))
                slope = float(sd[f"de.de_module.{i}.{bnm + 1}.weight"].item())
            else:
                bn, slope = None, 0.0
            we, wo, bias, fill = packing.pack_deconv_parity(w, b, bn)
            co = _DEC_CH[i][1]
            P[f"de{i}"] = (ConvWeights(we, co), ConvWeights(wo, co), bias, fill, slope)
        self._packed = P

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("dpcrn (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        assert x.dim() == 4 and x.shape[1] == 2 and x.shape[3] == self.N_BINS, tuple(x.shape)
        xn = x.float().permute(0, 2, 3, 1).contiguous()                  # [B,T,161,2]
        est = ops.cmul(xn, self.mask_nhwc(xn, taps))                    # DPCRN.py:33-42
        return est.permute(0, 3, 1, 2)

    def forward_nhwc(self, x, taps=None):
        """x [B,T,161,2] channels-last -> masked spectrum [B,T,161,2]."""
        return ops.cmul(x, self.mask_nhwc(x, taps))

    # ---- DPRNN ---------------------------------------------------------------------------------
    def _dprnn(self, x: Act, b, t):
        """x: fp32 + TF32 pair of [B,T,4,128]; returns the same (DPCRN.py:60-98)."""
        P = self._packed
        dev = x.f32.device
        m = b * t
        mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)   # noqa: E731
        # -- intra: Bi-LSTM(128 -> 2 x 64, 2 layers) over the 4 frequency positions of every frame
        src = (x.pair[0].view(m, 4, 128), x.pair[1].view(m, 4, 128))
        cst = mk(m, 64)
        for l in range(2):
            dst = (mk(m, 4, 128), mk(m, 4, 128))
            for d in range(2):
                w_hi, w_lo, bias = P[f"intra{l}{d}"]
                order = (0, 1, 2, 3) if d == 0 else (3, 2, 1, 0)
                prev = None
                for f in order:
                    h_hi, h_lo = dst[0][:, f, 64 * d:64 * d + 64], dst[1][:, f, 64 * d:64 * d + 64]
                    ops.lstm_cell_tf32x3_ex((src[0][:, f], src[1][:, f]), prev, w_hi, w_lo, bias, cst, h_hi, h_lo)
                    prev = (h_hi, h_lo)
            src = dst
        hi, lo, bias, _ = P["intra_fc"]
        y, _ = ops.gemm_tf32x3_ex((src[0].view(m * 4, 128), src[1].view(m * 4, 128)), hi, lo, bias, 128)
        g, be = P["ln1"]
        intra_f32, intra_pair = ops.group_layernorm(y.view(m, 512), 1, g, be, res=x.f32.view(m, 512), want_f32=True,
                                                    want_pair=True)
        # -- inter: LSTM(128 -> 128, 2 layers) over T for every (b, f); the 4 positions are the groups of one launch
        seq, pair = intra_f32.view(m * 4, 128), (intra_pair[0].view(m * 4, 128), intra_pair[1].view(m * 4, 128))
        for l in range(2):
            lay = P[f"inter{l}"]
            xp = lstm_engine.input_projection(seq, lay, pair)             # [B*T*4, 512]
            hs = mk(b, t, 4 * 128)
            ops.lstm_seq_multi(xp.view(b, t, 4 * 512), lay["whh"], 128, 4, hs)
            seq, pair = hs.view(m * 4, 128), None
        hi, lo, bias, w_kn = P["inter_fc"]
        if lstm_engine.USE_TENSOR_CORES:
            y, _ = ops.gemm_tf32x3_ex(ops.split_tf32(seq), hi, lo, bias, 128)
        else:
            y = ops.linear(seq, w_kn, bias, 128)
        g, be = P["ln2"]
        o_f32, o_pair = ops.group_layernorm(y.view(m, 512), 1, g, be, res=intra_f32, want_f32=True, want_pair=True)
        return Act(o_f32.view(b, t, 4, 128), (o_pair[0].view(b, t, 4, 128), o_pair[1].view(b, t, 4, 128)))

    def mask_nhwc(self, x, taps=None):
        """x [B,T,161,2] -> complex ratio mask [B,T,161,2] (everything of forward() but the final multiply)."""
        self._ensure_packed()
        P = self._packed
        b, t = x.shape[0], x.shape[1]
        dev = x.device
        tc = conv_engine.tc_eligible
        # ---- encoder ----
        enc = []
        h = Act(x)
        for i in range(5):
            w, bias, slope = P[f"en{i}"]
            ci, co = _ENC_CH[i], _ENC_CH[i + 1]
            # every activation keeps its fp32 copy (residuals, FMA layers); a tensor-core layer also emits the
            # TF32 split in its epilogue, an FMA layer's consumers split lazily (Act.get_pair)
            this_tc = tc(ci, 0, co, _ENC_F[i + 1], 2)
            out = conv_engine.new_act(b, t, _ENC_F[i + 1], co, dev, want_f32=True, want_pair=this_tc)
            conv_engine.conv(h, None, b, t, _ENC_F[i], _ENC_F[i + 1], packing.CONV23_TAPS, 2, w, bias, "prelu", out,
                             _ENC_F[i + 1], act_param=slope)
            h = out
            enc.append(h)
            if taps is not None:
                taps[f"en{i + 1}"] = h.f32
        # ---- DPRNN x 2, shared weights (DPCRN.py:28-29) ----
        if h.pair is None:
            h.get_pair()
        for r in range(2):
            h = self._dprnn(h, b, t)
            if taps is not None:
                taps[f"dp{r + 1}"] = h.f32
        # ---- decoder ----
        fin = 4
        for i in range(5):
            we, wo, bias, fill, slope = P[f"de{i}"]
            co = _DEC_CH[i][1]
            skip = enc[4 - i]
            shift = 1 if i == 3 else 0                                   # de4: left pad on F (DPCRN.py:159-165)
            fo = 2 * fin + 1 + shift
            c0, c1 = h.shape[-1], skip.shape[-1]
            is_tc = tc(c0, c1, co, fin + 1, 1)
            act = "prelu" if i < 4 else "none"
            out = conv_engine.new_act(b, t, fo, co, dev, want_f32=True, want_pair=is_tc and not shift)
            conv_engine.conv(h, skip, b, t, fin, fin + 1, packing.DECONV_EVEN_TAPS, 1, we, bias, act, out, fo,
                             dst_f0=shift, dst_fstep=2, act_param=slope)
            conv_engine.conv(h, skip, b, t, fin, fin, packing.DECONV_ODD_TAPS, 1, wo, bias, act, out, fo,
                             dst_f0=shift + 1, dst_fstep=2, act_param=slope, fill_f=(0 if shift else -1),
                             fill=(fill if shift else None))
            h = out
            fin = fo
            if taps is not None:
                taps[f"de{i + 1}"] = h.f32
        return h.f32
