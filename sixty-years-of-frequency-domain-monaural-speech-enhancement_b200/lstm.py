"""Drop-in ``lstm_net`` (reference: LSTM/LSTM.py:14-29) on the sm_100a kernels.

forward(x): [B,T,161] magnitude -> [B,T,161].  BatchNorm1d(161) (LSTM.py:16,25) is folded into
the first LSTM's input projection; Linear(1024,161)+Softplus (LSTM.py:19-22) is the GEMM
epilogue.  State-dict keys as shipped (LSTM/lstm_decode_vb.py:18-19).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lstm_engine, ops, packing
from .param_tree import bn_rows, build_param_tree, lstm_rows


def _spec():
    rows = bn_rows("bn", 161)
    rows += lstm_rows("lstm1", 161, 1024, 1)
    rows += lstm_rows("lstm2", 1024, 1024, 2)
    rows += [("fc.0.weight", (161, 1024), "param"), ("fc.0.bias", (161,), "param")]
    return rows


class lstm_net(nn.Module):
    N_BINS = 161

    def __init__(self):
        super().__init__()
        build_param_tree(self, _spec())
        self._packed = None
        self._packed_key = None

    def _state_key(self):
        p = next(self.parameters())
        return (p.device, tuple(int(t._version) for t in self.state_dict().values()))

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        s, o = packing.bn_fold(sd["bn.weight"], sd["bn.bias"], sd["bn.running_mean"], sd["bn.running_var"])
        P = {}
        P["l0"] = packing.pack_lstm_layer(sd["lstm1.weight_ih_l0"], sd["lstm1.weight_hh_l0"], sd["lstm1.bias_ih_l0"],
                                          sd["lstm1.bias_hh_l0"], in_scale=s, in_shift=o)
        for l in range(2):
            P[f"l{l + 1}"] = packing.pack_lstm_layer(sd[f"lstm2.weight_ih_l{l}"], sd[f"lstm2.weight_hh_l{l}"],
                                                     sd[f"lstm2.bias_ih_l{l}"], sd[f"lstm2.bias_hh_l{l}"])
        P["fc_w"] = packing.pad_cols(sd["fc.0.weight"].t().contiguous())
        P["fc_hi"], P["fc_lo"] = packing.split_tf32(sd["fc.0.weight"].contiguous())   # [161, 1024] K-major
        P["fc16"] = packing.pack_linear_f16(sd["fc.0.weight"].contiguous())
        P["fc_b"] = sd["fc.0.bias"].contiguous()
        self._packed = P

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def stream_cells(self):
        """The three LSTM layers packed for the one-step cell GEMM (streaming.MagStream); the eval BatchNorm1d in front of
        the first layer (LSTM.py:16,25) is folded into its input weights exactly as in the offline packing."""
        self._ensure_packed()
        if "cells" not in self._packed:
            sd = {k: v.detach().float() for k, v in self.state_dict().items()}
            s, o = packing.bn_fold(sd["bn.weight"], sd["bn.bias"], sd["bn.running_mean"], sd["bn.running_var"])
            w0 = sd["lstm1.weight_ih_l0"]
            cells = [packing.pack_lstm_cell((w0 * s[None, :]).contiguous(), sd["lstm1.weight_hh_l0"],
                                            sd["lstm1.bias_ih_l0"] + w0 @ o, sd["lstm1.bias_hh_l0"])]
            for l in range(2):
                cells.append(packing.pack_lstm_cell(sd[f"lstm2.weight_ih_l{l}"], sd[f"lstm2.weight_hh_l{l}"],
                                                    sd[f"lstm2.bias_ih_l{l}"], sd[f"lstm2.bias_hh_l{l}"]))
            self._packed["cells"] = cells
        return self._packed["cells"]

    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("lstm_net (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        self._ensure_packed()
        P = self._packed
        x = x.contiguous().float()
        b, t, f = x.shape
        assert f == self.N_BINS
        seq = x.view(b * t, f)
        for l in range(3):
            hs = lstm_engine.lstm_layer(seq, P[f"l{l}"], b, t)
            seq = hs.view(b * t, 1024)
            if taps is not None:
                taps[f"h{l}"] = hs
        if lstm_engine.USE_TENSOR_CORES and lstm_engine.USE_F16_PAIRS and seq.shape[0] >= 128:
            y = ops.gemm_f16x3(ops.split_f16(seq), P["fc16"][:2], P["fc16"][2], P["fc_b"], 161, act="softplus")
        elif lstm_engine.USE_TENSOR_CORES and seq.shape[0] >= 128:
            a_hi, a_lo = ops.split_tf32(seq)
            y = ops.gemm_tf32x3(a_hi, a_lo, P["fc_hi"], P["fc_lo"], P["fc_b"], 161, act="softplus")
        else:
            y = ops.linear(seq, P["fc_w"], P["fc_b"], 161, act="softplus")
        return y.view(b, t, 161)
