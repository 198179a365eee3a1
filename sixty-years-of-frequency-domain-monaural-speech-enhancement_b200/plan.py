"""ctypes front end of the plan-level C ABI (include/se_b200.h: se_plan_create_crn / se_forward_crn / se_enhance_crn /
se_query_workspace).  Deliberately imports NO model code (crn.py, packing.py, conv_engine.py ...): everything between the
reference state-dict and the enhanced waveform happens inside libse_b200.so (csrc/plan_crn.cu), exactly as a C host
would drive it -- see INTEGRATION.md for the equivalent C snippet.  torch is used for device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import CrnWeights, check


class CrnPlan:
    """crn_net + CRN/crn_decode.py:38-57 behind the plan ABI.  ``state_dict``: the reference checkpoint as torch.load returns
    it (CPU or CUDA tensors; copied to host fp32)."""

    def __init__(self, state_dict, b_max=64, n_max=64000, graph=False):
        self._host = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()
                      if v.dtype.is_floating_point}
        w = CrnWeights()
        ptr = lambda k: self._host[k].data_ptr()   # noqa: E731
        for i in range(5):
            w.en_w[i], w.en_b[i] = ptr(f"en.en_module.{i}.1.weight"), ptr(f"en.en_module.{i}.1.bias")
            bn = 3 if i == 3 else 2                # de4 carries an extra pad module (CRN/CRN.py:92-97)
            w.de_w[i], w.de_b[i] = ptr(f"de.de_module.{i}.0.weight"), ptr(f"de.de_module.{i}.0.bias")
            for j, n in enumerate(("weight", "bias", "running_mean", "running_var")):
                w.en_bn[i][j] = ptr(f"en.en_module.{i}.2.{n}")
                w.de_bn[i][j] = ptr(f"de.de_module.{i}.{bn}.{n}")
        for l in range(2):
            w.lstm_w_ih[l], w.lstm_w_hh[l] = ptr(f"lstm.weight_ih_l{l}"), ptr(f"lstm.weight_hh_l{l}")
            w.lstm_b_ih[l], w.lstm_b_hh[l] = ptr(f"lstm.bias_ih_l{l}"), ptr(f"lstm.bias_hh_l{l}")
        self._lib = _lib.load()
        self._plan = C.c_void_p()
        check(self._lib.se_plan_create_crn(C.byref(w), int(b_max), int(n_max), C.byref(self._plan)), "se_plan_create_crn")
        self.b_max, self.n_max = b_max, n_max
        if graph:
            check(self._lib.se_plan_set_graph(self._plan, 1), "se_plan_set_graph")

    @property
    def workspace_bytes(self):
        return int(self._lib.se_query_workspace(self._plan))

    def forward(self, mag):
        """[B,T,161] float32 CUDA -> [B,T,161]."""
        mag = mag.contiguous()
        est = torch.empty_like(mag)
        check(self._lib.se_forward_crn(self._plan, C.c_void_p(mag.data_ptr()), C.c_void_p(est.data_ptr()), mag.shape[0],
                                       mag.shape[1], C.c_void_p(torch.cuda.current_stream().cuda_stream)), "se_forward_crn")
        return est

    def enhance(self, wav, lengths=None, p=1.0):
        """[B,N] float32 CUDA -> enhanced [B,N]; ``lengths``: int32 CUDA [B] for a tail-padded batch."""
        wav = wav.contiguous()
        out = torch.empty_like(wav)
        lp = C.c_void_p(0 if lengths is None else lengths.data_ptr())
        check(self._lib.se_enhance_crn(self._plan, C.c_void_p(wav.data_ptr()), wav.stride(0), C.c_void_p(out.data_ptr()),
                                       out.stride(0), wav.shape[0], wav.shape[1], lp, float(p),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "se_enhance_crn")
        return out

    def close(self):
        if self._plan:
            self._lib.se_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
