"""Drop-in ``crn_net`` (reference: CRN/CRN.py:16-33) executing on the sm_100a kernels.

Same class name, constructor, ``forward(x)`` contract ([B,T,161] magnitude -> [B,T,161]
estimated magnitude) and state-dict keys as the reference, so
``crn_net().load_state_dict(torch.load('BEST_MODEL/wsj0_si84_300h_crn_noncprs_model.pth'))``
works unchanged (CRN/crn_decode.py:18-22).  Inference only.

Data layout (differs from the reference's NCHW on purpose): activations are channels-last
[B, T, F, C] so that (a) implicit-GEMM K-slices are contiguous, (b) the decoder's
``torch.cat((x, skip), 1)`` (CRN.py:107) is two pointers instead of a copy, and (c) the
encoder output [B,T,4,256] *is* the LSTM input [B,T,1024] -- the (c,f)->(f,c) flatten order
difference (CRN.py:27-28,31-32) is absorbed into the LSTM weight packing.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import conv_engine, lstm_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import bn_rows, build_param_tree, lstm_rows

_ENC_CH = [1, 16, 32, 64, 128, 256]          # CRN.py:40-62
_ENC_F = [161, 80, 39, 19, 9, 4]
_DEC_CH = [(512, 128), (256, 64), (128, 32), (64, 16), (32, 1)]   # CRN.py:77-99


def _spec():
    rows = []
    for i in range(5):
        ci, co = _ENC_CH[i], _ENC_CH[i + 1]
        rows += [(f"en.en_module.{i}.1.weight", (co, ci, 2, 3), "param"),
                 (f"en.en_module.{i}.1.bias", (co,), "param")]
        rows += bn_rows(f"en.en_module.{i}.2", co)
    rows += lstm_rows("lstm", 1024, 1024, 2)
    for i, (ci, co) in enumerate(_DEC_CH):
        bn = 3 if i == 3 else 2                   # de4 has the extra pad module (CRN.py:92-97)
        rows += [(f"de.de_module.{i}.0.weight", (ci, co, 2, 3), "param"),
                 (f"de.de_module.{i}.0.bias", (co,), "param")]
        rows += bn_rows(f"de.de_module.{i}.{bn}", co)
    return rows


class crn_net(nn.Module):
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

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        dev = next(self.parameters()).device
        P = {}
        for i in range(5):
            bn = tuple(sd[f"en.en_module.{i}.2.{n}"] for n in ("weight", "bias", "running_mean", "running_var"))
            w, bias = packing.pack_conv(sd[f"en.en_module.{i}.1.weight"], sd[f"en.en_module.{i}.1.bias"], bn)
            P[f"en{i}"] = (w, bias) if i == 0 else (ConvWeights(w, _ENC_CH[i + 1]), bias)
        # NHWC flatten index q = f*256 + c  <->  reference feature index c*4 + f
        q = torch.arange(1024, device=dev)
        nhwc = (q % 256) * 4 + q // 256
        P["lstm0"] = packing.pack_lstm_layer(sd["lstm.weight_ih_l0"], sd["lstm.weight_hh_l0"], sd["lstm.bias_ih_l0"],
                                             sd["lstm.bias_hh_l0"], in_perm=nhwc)
        P["lstm1"] = packing.pack_lstm_layer(sd["lstm.weight_ih_l1"], sd["lstm.weight_hh_l1"], sd["lstm.bias_ih_l1"],
                                             sd["lstm.bias_hh_l1"], unit_perm=nhwc)
        for i in range(5):
            bnm = 3 if i == 3 else 2
            bn = tuple(sd[f"de.de_module.{i}.{bnm}.{n}"] for n in ("weight", "bias", "running_mean", "running_var"))
            w, b = sd[f"de.de_module.{i}.0.weight"], sd[f"de.de_module.{i}.0.bias"]
            if i < 4:
                we, wo, bias, fill = packing.pack_deconv_parity(w, b, bn)
                co = _DEC_CH[i][1]
                cwe, cwo = ConvWeights(we, co), ConvWeights(wo, co)
                P[f"de{i}"] = (cwe, cwo, bias, fill)
                P[f"de{i}_m"] = conv_engine.merge_parity(cwe, packing.DECONV_EVEN_TAPS, cwo, packing.DECONV_ODD_TAPS)
            else:
                s, o = packing.bn_fold(*bn)
                wf = w * s[None, :, None, None]
                P["de4_w"] = wf[:, 0].permute(1, 2, 0).reshape(6, w.shape[0]).contiguous()   # [kt*3+kf][ci]
                P["de4_b"] = float((b * s + o).item())
        self._packed = P

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("crn_net (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        self._ensure_packed()
        P = self._packed
        x = x.contiguous().float()
        b, t, f = x.shape
        assert f == self.N_BINS, f"CRN checkpoints are hard-wired to 161 bins, got {f}"
        enc = self._encoder(x, taps)      # tensor-core layers on fp16 operand pairs unless SE_F16_PAIRS=0
        h = enc[-1]
        # LSTM: two layers, input projection hoisted over all T
        seq, pair = (h.f32.view(b * t, 1024) if h.f32 is not None else None,
                     (h.pair[0].view(b * t, 1024), h.pair[1].view(b * t, 1024)) if h.pair is not None else None)
        for l in range(2):
            hs = lstm_engine.lstm_layer(seq, P[f"lstm{l}"], b, t, pair)
            seq, pair = hs.view(b * t, 1024), None
        if taps is not None:
            taps["lstm_nhwc"] = hs
        return self._decoder(hs, enc, taps)

    def _encoder(self, x, taps=None, f16=None):
        """CRN/CRN.py:35-71 on a [B,T,161] plane -> the five channels-last encoder activations (the last one, [B,T,4,256],
        is the LSTM input [B,T,1024]).  f16: operand pairs of the tensor-core layers (None = lstm_engine.USE_F16_PAIRS)."""
        P = self._packed
        b, t, _ = x.shape
        dev = x.device
        f16 = lstm_engine.USE_F16_PAIRS if f16 is None else f16
        tc = lambda c0, c1, co, fo, sf: conv_engine.tc_eligible(c0, c1, co, fo, sf, f16)   # noqa: E731
        enc = []
        w, bias = P["en0"]
        h = Act(ops.conv_in1(x, w, bias, 16, "elu", _ENC_F[1]))
        enc.append(h)
        for i in range(1, 5):
            w, bias = P[f"en{i}"]
            ci, co = _ENC_CH[i], _ENC_CH[i + 1]
            is_tc = tc(ci, 0, co, _ENC_F[i + 1], 2)
            # every consumer of en2..en5 (next encoder layer / LSTM projection / decoder skip) is a
            # tensor-core layer when tensor cores are on: emit the TF32 split only
            out = conv_engine.new_act(b, t, _ENC_F[i + 1], co, dev, want_f32=not is_tc, want_pair=is_tc, f16=f16)
            conv_engine.conv(h, None, b, t, _ENC_F[i], _ENC_F[i + 1], packing.CONV23_TAPS, 2, w, bias, "elu", out,
                             _ENC_F[i + 1])
            h = out
            enc.append(h)
        if taps is not None:
            for i, e in enumerate(enc):
                taps[f"en{i + 1}"] = e.value()
        return enc

    def _decoder(self, hs, enc, taps=None, f16=None):
        """CRN/CRN.py:73-109: hs [B,T,1024] (NHWC-flattened LSTM output) + encoder skips -> [B,T,161]."""
        P = self._packed
        b, t = hs.shape[0], hs.shape[1]
        dev = hs.device
        f16 = lstm_engine.USE_F16_PAIRS if f16 is None else f16
        tc = lambda c0, c1, co, fo, sf: conv_engine.tc_eligible(c0, c1, co, fo, sf, f16)   # noqa: E731
        h = Act(hs.view(b, t, 4, 256))
        fin = 4
        for i in range(4):
            we, wo, bias, fill = P[f"de{i}"]
            co = _DEC_CH[i][1]
            skip = enc[4 - i]
            shift = 1 if i == 3 else 0                      # de4: left pad on F (CRN.py:92-97)
            fo = 2 * fin + 1 + shift
            c0, c1 = h.shape[-1], skip.shape[-1]
            is_tc = tc(c0, c1, co, fin + 1, 1)
            last = i == 3                                   # de4 feeds the fp32 direct kernel
            out = conv_engine.new_act(b, t, fo, co, dev, want_f32=(not is_tc) or last, want_pair=is_tc and not last,
                                      f16=f16)
            wm = P[f"de{i}_m"]
            if f16 and conv_engine.parity2_eligible(h, skip, wm, fin + 1):
                # both output-column parity classes in one launch: the activation tiles are read once
                conv_engine.conv_parity2(h, skip, b, t, fin, fin + 1, fin, packing.DECONV_EVEN_TAPS, wm, bias, "elu", out,
                                         fo, dst_f0=shift)
                if shift:
                    ops.fill_column(out.f32, fill, 0, "elu", 0.0)
            else:
                conv_engine.conv(h, skip, b, t, fin, fin + 1, packing.DECONV_EVEN_TAPS, 1, we, bias, "elu", out, fo,
                                 dst_f0=shift, dst_fstep=2)
                conv_engine.conv(h, skip, b, t, fin, fin, packing.DECONV_ODD_TAPS, 1, wo, bias, "elu", out, fo,
                                 dst_f0=shift + 1, dst_fstep=2, fill_f=(0 if shift else -1),
                                 fill=(fill if shift else None))
            h = out
            fin = fo
            if taps is not None:
                taps[f"de{i + 1}"] = h.value()
        h = h.f32
        enc0 = enc[0].f32
        y = ops.deconv_out1(h, enc0, P["de4_w"], P["de4_b"], "softplus")
        return y

    # -- streaming ---------------------------------------------------------------
    STREAM_CONTEXT = (5, 5)      # frames of past the encoder / decoder conv stacks look at (one per k(2,.) layer)

    def stream_cells(self):
        """The two LSTM layers packed for the one-step cell GEMM (se_lstm_cell_tf32x3_ex) with the same NHWC
        permutations as the offline packing: layer 0 reads the NHWC-flattened encoder output, layer 1 emits NHWC."""
        self._ensure_packed()
        if "cells" not in self._packed:
            sd = {k: v.detach().float() for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            q = torch.arange(1024, device=dev)
            nhwc = (q % 256) * 4 + q // 256
            rows = (torch.arange(4, device=dev).view(4, 1) * 1024 + nhwc.view(1, -1)).reshape(-1)
            c0 = packing.pack_lstm_cell(sd["lstm.weight_ih_l0"][:, nhwc].contiguous(), sd["lstm.weight_hh_l0"],
                                        sd["lstm.bias_ih_l0"], sd["lstm.bias_hh_l0"])
            c1 = packing.pack_lstm_cell(sd["lstm.weight_ih_l1"][rows].contiguous(),
                                        sd["lstm.weight_hh_l1"][rows][:, nhwc].contiguous(),
                                        sd["lstm.bias_ih_l1"][rows], sd["lstm.bias_hh_l1"][rows])
            self._packed["cells"] = [c0, c1]
        return self._packed["cells"]
