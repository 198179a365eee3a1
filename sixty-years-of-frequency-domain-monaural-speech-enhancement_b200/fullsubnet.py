"""Drop-in FullSubNet ``Model`` (reference: FullSubNet/fullsubnet_net_sa/model.py:9-118).

Same constructor arguments, ``forward(noisy_mag [B,1,257,T]) -> complex mask [B,2,257,T]`` and
state-dict keys (``fb_model.sequence_model.weight_ih_l0`` ...,
FullSubNet/fullsubnet_sa_decode.py:11-27).  Semantics are PER UTTERANCE for every batch size:
the reference decodes one file at a time, where ``if batch_size > 1: drop_band`` (model.py:101-104,
a training-time trick that halves the frequency axis) never runs -- see SURVEY.md section 0.1.

Execution plan (B clips, T frames, Tp = T + look_ahead):
  full-band model: [B,Tp,257] -> two LSTM(512) layers through the persistent recurrence kernel
                   (H = 512: 64 weight-stationary CTAs) -> Linear(512,257)+ReLU on tensor cores;
  sub-band model : B*257 independent sequences of 32 features.  Its hidden sequence
                   ([B*257, Tp, 384] = 7.9 GB per layer at config 4) is never materialised: the
                   two layers advance together one time step at a time, each step one fused
                   tcgen05 3xTF32 GEMM [B*257, 32+384 | 384+384] x [., 1536] whose epilogue is
                   the LSTM cell; Linear(384,2) is applied to the step's hidden state.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lstm_engine, ops, packing
from .param_tree import build_param_tree, lstm_rows


def _spec(num_freqs, fb_hidden, sb_hidden, sb_in):
    rows = lstm_rows("fb_model.sequence_model", num_freqs, fb_hidden, 2)
    rows += [("fb_model.fc_output_layer.weight", (num_freqs, fb_hidden), "param"),
             ("fb_model.fc_output_layer.bias", (num_freqs,), "param")]
    rows += lstm_rows("sb_model.sequence_model", sb_in, sb_hidden, 2)
    rows += [("sb_model.fc_output_layer.weight", (2, sb_hidden), "param"),
             ("sb_model.fc_output_layer.bias", (2,), "param")]
    return rows


class Model(nn.Module):
    def __init__(self, num_freqs, look_ahead, sequence_model, fb_num_neighbors, sb_num_neighbors,
                 fb_output_activate_function, sb_output_activate_function, fb_model_hidden_size,
                 sb_model_hidden_size, norm_type="offline_laplace_norm", num_groups_in_drop_band=2,
                 weight_init=True):
        super().__init__()
        if sequence_model != "LSTM" or fb_num_neighbors != 0 or norm_type != "offline_laplace_norm":
            raise NotImplementedError("se_b200 FullSubNet covers the configuration the decode scripts use: "
                                      "LSTM, fb_num_neighbors=0, offline_laplace_norm")
        if fb_output_activate_function != "ReLU" or sb_output_activate_function is not None:
            raise NotImplementedError("fb ReLU / sb linear outputs only (fullsubnet_sa_decode.py:17-18)")
        self.num_freqs, self.look_ahead = num_freqs, look_ahead
        self.sb_num_neighbors = sb_num_neighbors
        self.fb_hidden, self.sb_hidden = fb_model_hidden_size, sb_model_hidden_size
        self.sb_in = (sb_num_neighbors * 2 + 1) + 1
        build_param_tree(self, _spec(num_freqs, self.fb_hidden, self.sb_hidden, self.sb_in))
        self._packed = None
        self._packed_key = None

    def _state_key(self):
        p = next(self.parameters())
        return (p.device, tuple(int(t._version) for t in self.state_dict().values()))

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        dev = next(self.parameters()).device
        P = {}
        for l in range(2):
            pre = "fb_model.sequence_model"
            P[f"fb{l}"] = packing.pack_lstm_layer(sd[f"{pre}.weight_ih_l{l}"], sd[f"{pre}.weight_hh_l{l}"],
                                                  sd[f"{pre}.bias_ih_l{l}"], sd[f"{pre}.bias_hh_l{l}"])
            pre = "sb_model.sequence_model"
            P[f"sb{l}"] = packing.pack_lstm_cell(sd[f"{pre}.weight_ih_l{l}"], sd[f"{pre}.weight_hh_l{l}"],
                                                 sd[f"{pre}.bias_ih_l{l}"], sd[f"{pre}.bias_hh_l{l}"])
            P[f"sb{l}_f16"] = packing.pack_lstm_cell_f16(sd[f"{pre}.weight_ih_l{l}"], sd[f"{pre}.weight_hh_l{l}"],
                                                         sd[f"{pre}.bias_ih_l{l}"], sd[f"{pre}.bias_hh_l{l}"])
        if self.fb_hidden == 512:
            # H = 512 as the first block of a block-diagonal H = 1024 recurrence: the tcgen05 cluster kernel (lstm_f16.cu)
            # exists for H = 1024 only, and its 5.4 us per step beat the 8.9 us of the fp32 slice kernel at H = 512 even with
            # half of the units idle (their xproj columns and weights are zero: gates 0.5 / 0, c = h = 0)
            for l in range(2):
                bd = torch.zeros(128, 1024, 32, device=dev)
                bd[:64, :512] = P[f"fb{l}"]["whh"]
                P[f"fb{l}_bd"] = bd.contiguous()
        w = sd["fb_model.fc_output_layer.weight"].contiguous()
        P["fb_fc16"] = packing.pack_linear_f16(w)
        P["fb_fc_hi"], P["fb_fc_lo"] = packing.split_tf32(w)
        P["fb_fc_kn"] = packing.pad_cols(w.t().contiguous())
        P["fb_fc_b"] = sd["fb_model.fc_output_layer.bias"].contiguous()
        P["sb_fc_w"] = sd["sb_model.fc_output_layer.weight"].contiguous()
        P["sb_fc_b"] = sd["sb_model.fc_output_layer.bias"].contiguous()
        # how often each magnitude bin appears in the reflect-padded unfold (for the sb norm mean)
        f, n = self.num_freqs, self.sb_num_neighbors
        idx = torch.arange(f)[:, None] + torch.arange(2 * n + 1)[None, :] - n
        idx = torch.where(idx < 0, -idx, idx)
        idx = torch.where(idx >= f, 2 * (f - 1) - idx, idx)
        P["unfold_count"] = torch.bincount(idx.reshape(-1), minlength=f).float().to(dev)
        self._packed = P

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, noisy_mag, taps=None):
        if not noisy_mag.is_cuda:
            raise RuntimeError("FullSubNet (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(noisy_mag, taps)

    def _forward_impl(self, noisy_mag, taps=None):
        assert noisy_mag.dim() == 4 and noisy_mag.shape[1] == 1 and noisy_mag.shape[2] == self.num_freqs
        self._ensure_packed()
        P = self._packed
        x = noisy_mag.float()
        b, _, f, t = x.shape
        tp = t + self.look_ahead
        dev = x.device
        strides = (x.stride(0), x.stride(3), x.stride(2))          # (sb, st, sf) of the [B,T,F] view

        # ---- full-band model ------------------------------------------------------------------
        inv_fb = ops.fsn_clip_inv_mean(x, strides, b, t, f, denom=float(f * tp))
        mag_tm, xn = ops.fsn_fb_input(x, strides, b, t, tp, f, inv_fb)
        seq = xn.view(b * tp, f)
        f16 = lstm_engine.USE_TENSOR_CORES and lstm_engine.USE_F16_PAIRS and b * tp >= 128
        for l in range(2):
            if f16 and f"fb{l}_bd" in P and b <= 64:
                lay = P[f"fb{l}"]
                xp = torch.zeros(b * tp, 4096, device=dev, dtype=torch.float32)          # columns of the idle units stay 0
                ops.gemm_f16x3(ops.split_f16(seq), (lay["wih16_hi"], lay["wih16_lo"]), lay["wih16_scale"], lay["bias"], 2048,
                               out=xp[:, :2048])
                hs = ops.lstm_seq(xp.view(b, tp, 4096), P[f"fb{l}_bd"], 1024)
                seq = hs.view(b * tp, 1024)[:, :512]                                       # strided view, no copy
            else:
                hs = lstm_engine.lstm_layer(seq, P[f"fb{l}"], b, tp)
                seq = hs.view(b * tp, self.fb_hidden)
        if f16:
            fb_out = ops.gemm_f16x3(ops.split_f16(seq), P["fb_fc16"][:2], P["fb_fc16"][2], P["fb_fc_b"], f, act="relu")
        elif lstm_engine.USE_TENSOR_CORES and seq.shape[0] >= 128:
            a_hi, a_lo = ops.split_tf32(seq)
            fb_out = ops.gemm_tf32x3(a_hi, a_lo, P["fb_fc_hi"], P["fb_fc_lo"], P["fb_fc_b"], f, act="relu")
        else:
            fb_out = ops.linear(seq, P["fb_fc_kn"], P["fb_fc_b"], f, act="relu")
        fb_out = fb_out.view(b, tp, f)
        if taps is not None:
            taps["fb_out_btf"] = fb_out

        # ---- sub-band model -------------------------------------------------------------------
        nn_ = self.sb_num_neighbors
        inv_sb = ops.fsn_clip_inv_mean(mag_tm, (tp * f, f, 1), b, tp, f, denom=float(f * self.sb_in * tp),
                                       wgt=P["unfold_count"], extra=fb_out)
        m, hd = b * f, self.sb_hidden
        z = lambda: torch.zeros(m, hd, device=dev, dtype=torch.float32)   # noqa: E731
        if lstm_engine.USE_F16_PAIRS:
            # fp16 operand pairs: x, h and W travel as scaled fp16 (hi, lo); half the MMAs and bytes of the TF32 pairs
            sbx_hi, sbx_lo = ops.fsn_sb_assemble_f16(mag_tm, fb_out, nn_, inv_sb)      # [Tp, B*F, 32] fp16
            L0, L1 = P["sb0_f16"], P["sb1_f16"]
            z16 = lambda: torch.zeros(m, hd, device=dev, dtype=torch.float16)   # noqa: E731
            h0 = [(z16(), z16()), (z16(), z16())]
            h1 = [(z16(), z16()), (z16(), z16())]
            c0, c1 = z(), z()
            h1_out = torch.empty(m, hd, device=dev, dtype=torch.float32)
            mask = torch.empty(tp, m, 2, device=dev, dtype=torch.float32)
            for s in range(tp):
                src0, dst0 = h0[s & 1], h0[(s + 1) & 1]
                src1, dst1 = h1[s & 1], h1[(s + 1) & 1]
                ops.lstm_cell_f16x3((sbx_hi[s], sbx_lo[s]), src0, L0, c0, dst0[0], dst0[1])
                ops.lstm_cell_f16x3(dst0, src1, L1, c1, dst1[0], dst1[1], h1_out)
                ops.fsn_sb_fc(h1_out, P["sb_fc_w"], P["sb_fc_b"], mask[s])
            return mask.view(tp, b, f, 2).permute(1, 3, 2, 0)[:, :, :, self.look_ahead:]
        sbx_hi, sbx_lo = ops.fsn_sb_assemble(mag_tm, fb_out, nn_, inv_sb)      # [Tp, B*F, 32]
        L0, L1 = P["sb0"], P["sb1"]
        h0 = [(z(), z()), (z(), z())]      # layer-0 state (hi, lo), double buffered by step parity
        h1 = [(z(), z()), (z(), z())]
        c0, c1 = z(), z()
        h1_out = torch.empty(m, hd, device=dev, dtype=torch.float32)
        mask = torch.empty(tp, m, 2, device=dev, dtype=torch.float32)
        for s in range(tp):
            src0, dst0 = h0[s & 1], h0[(s + 1) & 1]
            src1, dst1 = h1[s & 1], h1[(s + 1) & 1]
            ops.lstm_cell_tf32x3(sbx_hi[s], sbx_lo[s], src0[0], src0[1], L0["w_hi"], L0["w_lo"], L0["bias"], c0,
                                 dst0[0], dst0[1])
            ops.lstm_cell_tf32x3(dst0[0], dst0[1], src1[0], src1[1], L1["w_hi"], L1["w_lo"], L1["bias"], c1,
                                 dst1[0], dst1[1], h1_out)
            ops.fsn_sb_fc(h1_out, P["sb_fc_w"], P["sb_fc_b"], mask[s])
        # [Tp, B, F, 2] -> the reference's [B, 2, F, T] (a view; look-ahead frames dropped, model.py:117)
        out = mask.view(tp, b, f, 2).permute(1, 3, 2, 0)[:, :, :, self.look_ahead:]
        return out


FullSubNet = Model
