"""Uformer: run the UNMODIFIED reference module (Uformer/uformer.py) on CPU -- build container only.

TEST INFRASTRUCTURE.  The reference calls the pre-1.8 ``torch.stft/istft`` real-view API and
``.cuda()`` unconditionally (uformer.py:178-186,276); both are shimmed here.
Used by oracle/make_golden.py to write tests/golden/uformer_*.npz (network taps + waveforms).
The travelling restatement is oracle.nets.uformer_forward / oracle.decode.enhance_uformer; this module pins
it (make_golden records the waveform max-abs difference in ``ref_vs_oracle``).
"""
from __future__ import annotations

import contextlib

import torch

from . import ref_shims


@contextlib.contextmanager
def legacy_torch_api():
    o_stft, o_istft, o_cuda = torch.stft, torch.istft, torch.Tensor.cuda

    def stft_legacy(x, n_fft, hop_length=None, win_length=None, window=None, **kw):
        return torch.view_as_real(o_stft(x, n_fft, hop_length=hop_length, win_length=win_length, window=window,
                                         return_complex=True))

    def istft_legacy(x, n_fft, hop_length=None, win_length=None, window=None, center=True, **kw):
        if not torch.is_complex(x):
            x = torch.view_as_complex(x.contiguous())
        return o_istft(x, n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center, **kw)

    torch.stft, torch.istft = stft_legacy, istft_legacy
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.stft, torch.istft, torch.Tensor.cuda = o_stft, o_istft, o_cuda


def build(sd):
    with legacy_torch_api():
        mod = ref_shims.import_reference("Uformer", "uformer")
        net = mod.Uformer().eval()
    net.load_state_dict(sd)
    return net


def run(net, wav, taps=None):
    """wav [B,N] float32 tensor -> (enhanced [B,hop*(T-1)], est complex [B,2,257,T]).  ``taps`` (dict) receives
    the outputs of the conformer sub-modules and the fused encoder levels."""
    hooks = []
    if taps is not None:
        c = net.conformer
        named = {"ff1_c": c.ff1_cplx, "ff1_m": c.ff1_mag, "tatt_c": c.cplx_tatt, "tatt_m": c.mag_tatt,
                 "fatt_c": c.cplx_fatt, "fatt_m": c.mag_fatt, "ff2_c": c.ff2_cplx, "ff2_m": c.ff2_mag}
        for i in (0, 3, 7):
            named[f"ds{i}_c"] = c.dsconv_cplx[i]
            named[f"ds{i}_m"] = c.dsconv_real[i]
        for i in range(6):
            named[f"encraw{i}_c"] = net.encoder[i]
            named[f"encraw{i}_m"] = net.encoder_real[i]
            named[f"decraw{i}_c"] = net.decoder[i]
            named[f"decraw{i}_m"] = net.decoder_real[i]
        for name, m in named.items():
            hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: taps.__setitem__(name, out.detach())))
        def conf_hook(mod, inp, out):
            taps["conf_c"], taps["conf_m"] = out[0].detach(), out[1].detach()
        hooks.append(c.register_forward_hook(conf_hook))
    with legacy_torch_api(), torch.no_grad():
        out = net(wav, wav)
    for h in hooks:
        h.remove()
    return out[0], out[2]
