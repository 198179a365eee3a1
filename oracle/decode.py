"""Oracle decode loops: restatements of the reference ``enhance()`` bodies, one clip at a time.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Waveform in (float64, as
``soundfile.read`` returns it) -> enhanced waveform out (float32), no disk I/O.
Every function returns ``(enhanced, taps)`` where ``taps`` holds the intermediate
arrays the parity tests bisect on, all on the c-normalised scale except ``wav``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import dsp, nets


def _mag_mapping_320(sd, forward, wav, p):
    """Shared body of ``CRN/crn_decode.py:38-57`` and ``LSTM/lstm_decode_vb.py:35-52``
    (librosa dialect, magnitude mapping with the noisy phase, backend rule (i))."""
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    x, c = dsp.rms_scale(wav)                                   # crn_decode.py:39-40
    spec = dsp.stft(x, n_fft, win, hop).T                       # :41   [T,161] complex64
    mag = np.abs(spec) ** p                                     # :44
    phase = np.angle(spec)
    feat = torch.from_numpy(mag.astype(np.float32))             # :46 FloatTensor
    with torch.no_grad():
        est = forward(sd, feat.unsqueeze(0)).squeeze(0).numpy() # :48
    est = est ** (1.0 / p)                                      # :51
    de = est * np.exp(1j * phase)                               # :54
    y = dsp.istft(de.T, n_fft, win, hop, length=len(x))         # :55-56
    y = y / c                                                   # :57
    taps = {"c": c, "mag": mag.astype(np.float32), "spec": spec, "est": est,
            "y_norm": (y * c).astype(np.float32)}
    return y.astype(np.float32), taps


def enhance_crn(sd, wav, p=1.0):
    """``CRN/crn_decode.py`` (p = 1.0 as shipped: ``**1.0`` at :44 and ``**1.`` at :51)."""
    return _mag_mapping_320(sd, nets.crn_forward, wav, p)


def enhance_lstm(sd, wav, p=1.0):
    """``LSTM/lstm_decode_vb.py`` at 16 kHz (the 48k->16k resample at :34 sits before the path)."""
    return _mag_mapping_320(sd, nets.lstm_net_forward, wav, p)


def dsp_identity(wav, geometry="fullsubnet"):
    """STFT -> identity mask -> iSTFT at a given geometry (the literal 512/256 DSP-only run of
    SURVEY.md section 8(d)); returns the reconstructed waveform (float32)."""
    n_fft, win, hop = dsp.GEOMETRIES[geometry]
    x, c = dsp.rms_scale(wav)
    spec = dsp.stft(x.astype(np.float32), n_fft, win, hop)
    y = dsp.istft(spec, n_fft, win, hop, length=len(x))
    return (y / c).astype(np.float32)


def enhance_fullsubnet(sd, wav, p=0.5):
    """``FullSubNet/fullsubnet_sa_decode.py:44-78`` (torch dialect, complex mask applied in the
    script on the COMPRESSED spectrum, backend rule (iii); p = 0.5 for the cprs checkpoint the
    script loads at :25, 1.0 for the noncprs twin)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["fullsubnet"]
    x, c = dsp.rms_scale(wav)                                               # :46-47
    x32 = x.astype(np.float32)                                              # :52 FloatTensor
    spec = dsp.stft(x32, n_fft, win, hop)                                   # :53  [F,T] complex64
    mag = np.abs(spec) ** p                                                 # :57
    ph = np.angle(spec)                                                     # :58
    xr, xi = mag * np.cos(ph), mag * np.sin(ph)                             # :59
    feat_mag = np.sqrt(xr ** 2 + xi ** 2).astype(np.float32)                # :61
    with torch.no_grad():
        m = _n.fullsubnet_forward(sd, torch.from_numpy(feat_mag)[None, None]).squeeze(0).numpy()   # :63
    er = m[0] * xr - m[1] * xi                                              # :66
    ei = m[0] * xi + m[1] * xr                                              # :67
    emag = np.sqrt(er ** 2 + ei ** 2) ** (1.0 / p)                          # :71
    eph = np.arctan2(ei, er)                                                # :72
    est = emag * np.cos(eph) + 1j * emag * np.sin(eph)                      # :73
    y = dsp.istft(est.astype(np.complex64), n_fft, win, hop, length=len(x)) # :76
    y = y / c                                                               # :78
    taps = {"c": c, "mag": feat_mag, "mask": m, "y_norm": (y * c).astype(np.float32)}
    return y.astype(np.float32), taps


def enhance_dccrn(sd, wav, p=0.5, crop_first=True):
    """``DCCRN/dccrn_decode.py:30-60`` (torch dialect; zero-pad to a whole number of hops, compress,
    DCCRN-E forward, decompress (rule (ii)), torch.istft without ``length`` then ``[:wav_len]``).
    ``p=1.0, crop_first=False`` is ``DCCRN_SNR/dccrn_decode_snr.py:30-64`` (same loop, exponent 1., and the
    network of DCCRN_SNR/DCCRN.py whose decoder crops ``[..., :-1]``, :159)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["dccrn"]
    x, c = dsp.rms_scale(wav)                                               # :31-32
    wav_len = len(x)
    frame_num = int(np.ceil((wav_len - 512 + 512) / 128 + 1))               # :36
    fake_len = (frame_num - 1) * 128 + 512 - 512
    x32 = np.concatenate((x, np.zeros(fake_len - wav_len))).astype(np.float32)   # :39
    spec = dsp.stft(x32, n_fft, win, hop)                                   # :41 [F,T]
    mag, ph = np.abs(spec) ** p, np.angle(spec)                             # :44
    feat = np.stack((mag * np.cos(ph), mag * np.sin(ph))).astype(np.float32)     # :46 [2,F,T]
    with torch.no_grad():
        est = _n.dccrn_forward(sd, torch.from_numpy(feat)[None], crop_first=crop_first).squeeze(0).numpy()   # :48
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)                  # :49,52
    eph = np.arctan2(est[1], est[0])                                        # :50
    y = dsp.istft((emag * np.cos(eph) + 1j * emag * np.sin(eph)).astype(np.complex64), n_fft, win, hop, None)  # :56
    y = y[:wav_len]                                                         # :59
    y = y / c
    taps = {"c": c, "feat": feat, "est": est, "y_norm": (y * c).astype(np.float32)}
    return y.astype(np.float32), taps


def enhance_uformer_ref(net, wav):
    """``Uformer/uformer_decode.py:38-50`` around the UNMODIFIED reference module (oracle.uformer_ref;
    build container only): c-normalise (float64), FloatTensor, model(x, x)[0], / c."""
    from . import uformer_ref
    x, c = dsp.rms_scale(wav)                                               # :40-41
    y, est = uformer_ref.run(net, torch.from_numpy(x.astype(np.float32))[None])   # :43-45
    y = y.squeeze(0).numpy()
    taps = {"c": c, "est": est.squeeze(0).numpy(), "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps                                 # :47-48


def enhance_uformer(sd, wav):
    """``Uformer/uformer_decode.py:38-50`` with the functional restatement of the network
    (oracle.nets.uformer_forward); STFT / iSTFT as uformer.py:178,276 (512/400/160, no ``length``)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["uformer"]
    x, c = dsp.rms_scale(wav)
    spec = dsp.stft(x.astype(np.float32), n_fft, win, hop)                  # [F,T] complex64
    with torch.no_grad():
        er, ei = _n.uformer_forward(sd, torch.from_numpy(spec.real.copy())[None], torch.from_numpy(spec.imag.copy())[None])
    est = (er[0].numpy() + 1j * ei[0].numpy()).astype(np.complex64)
    y = dsp.istft(est, n_fft, win, hop, None)
    taps = {"c": c, "est": np.stack([est.real, est.imag]), "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps


def enhance_gcrn(sd, wav, p=0.5):
    """``GCRN/gcrn_decode_vb.py:34-58`` (librosa dialect, compressed RI in, RI out, backend rule (ii);
    p = 0.5 in the vb script (:40,51), 1.0 in gcrn_decode.py)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    x, c = dsp.rms_scale(wav)
    spec = dsp.stft(x, n_fft, win, hop).T                                   # [T,161] complex64
    mag, ph = (np.abs(spec) ** p).astype(np.float32), np.angle(spec).astype(np.float32)
    feat = np.stack((mag * np.cos(ph), mag * np.sin(ph)))                   # [2,T,161]  (:45)
    with torch.no_grad():
        est = _n.gcrn_forward(sd, torch.from_numpy(feat)[None]).squeeze(0).numpy()
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)                  # :48,51
    eph = np.arctan2(est[1], est[0])                                        # :49
    de = emag * np.exp(1j * eph)                                            # :55
    y = dsp.istft(de.T, n_fft, win, hop, length=len(x))                     # :56-57
    taps = {"c": c, "feat": feat, "est": est, "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps


def enhance_dpcrn(sd, wav, p=1.0):
    """``DPCRN/dpcrn_decode_vb.py:33-60`` (librosa dialect; p = 1.0 in the vb script (:41,53), 0.5 in
    drcrn_decode.py:45,56): compressed RI in, masked RI out (the mask multiply is inside forward),
    backend rule (ii)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    x, c = dsp.rms_scale(wav)
    spec = dsp.stft(x, n_fft, win, hop).T
    mag, ph = (np.abs(spec) ** p).astype(np.float32), np.angle(spec).astype(np.float32)
    feat = np.stack((mag * np.cos(ph), mag * np.sin(ph)))
    with torch.no_grad():
        est = _n.dpcrn_forward(sd, torch.from_numpy(feat)[None]).squeeze(0).numpy()
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)
    eph = np.arctan2(est[1], est[0])
    de = emag * np.exp(1j * eph)
    y = dsp.istft(de.T, n_fft, win, hop, length=len(x))
    taps = {"c": c, "feat": feat, "est": est, "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps


def enhance_ctsnet(sds, wav, p=1.0, cumulative=False):
    """``CTSNet/two_stage_com_decode_vb.py:61-95`` (torch dialect; p = 1.0 there (:73,87), 0.5 for the cprs
    checkpoints of two_stage_com_decode.py / CTSNet_new).  ``sds`` = (stage-1 state-dict, stage-2 state-dict).
    Stage 1 maps the magnitude, its estimate takes the noisy phase, stage 2 sees cat(noisy RI, stage-1 RI) and
    predicts a complex residual (backend rules (iv) + (ii)); iSTFT without ``length`` then ``[:wav_len]``."""
    from . import nets as _n
    sd1, sd2 = sds
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    x, c = dsp.rms_scale(wav)                                               # :62-63
    wav_len = len(x)
    frame_num = int(np.ceil((wav_len - 320 + 320) / 160 + 1))               # :65
    fake_len = (frame_num - 1) * 160 + 320 - 320
    x32 = np.concatenate((x, np.zeros(fake_len - wav_len))).astype(np.float32)   # :68 FloatTensor
    spec = dsp.stft(x32, n_fft, win, hop).T                                 # :69-70  [T,F] complex64
    mag, ph = (np.abs(spec) ** p).astype(np.float32), np.angle(spec).astype(np.float32)   # :73
    feat = np.stack((mag * np.cos(ph), mag * np.sin(ph))).astype(np.float32)             # :75  [2,T,F]
    with torch.no_grad():
        fx = torch.from_numpy(feat)[None]
        phase = torch.from_numpy(ph)[None]
        est1 = _n.ctsnet_step1_forward(sd1, torch.norm(fx, dim=1), cumulative)            # :79
        s1 = torch.stack((est1 * torch.cos(phase), est1 * torch.sin(phase)), dim=1)       # :80-81
        s2 = _n.ctsnet_step2_forward(sd2, torch.cat((fx, s1), dim=1), cumulative=cumulative) + s1   # :82-84
    est = s2.squeeze(0).numpy()
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)                  # :87
    eph = np.arctan2(est[1], est[0])                                        # :89
    y = dsp.istft((emag * np.cos(eph) + 1j * emag * np.sin(eph)).T.astype(np.complex64), n_fft, win, hop, None)  # :93
    y = y[:wav_len]                                                         # :95
    taps = {"c": c, "feat": feat, "est1": est1.squeeze(0).numpy(), "est": est, "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps


def enhance_taylorsenet(sd, wav, p=1.0, cumulative=False):
    """``TaylorSENet/taylorsenet_decode_vb.py:27-52`` (torch dialect; p = 1.0 there (:40,44), 0.5 in TaylorSENet_new and
    for the cprs checkpoints): zero-pad to whole hops, STFT, compressed RI in, RI out (the Taylor recursion is inside
    forward), decompress (rule (ii)), iSTFT(length=N)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    x, c = dsp.rms_scale(wav)                                               # :29-30
    wav_len = len(x)
    frame_num = int(np.ceil((wav_len - 320 + 320) / 160 + 1))               # :32
    fake_len = (frame_num - 1) * 160 + 320 - 320
    x32 = np.concatenate((x, np.zeros(fake_len - wav_len))).astype(np.float32)   # :35
    spec = dsp.stft(x32, n_fft, win, hop).T                                 # :36-37  [T,F]
    mag, ph = (np.abs(spec) ** p).astype(np.float32), np.angle(spec).astype(np.float32)   # :40
    feat = np.stack((mag * np.cos(ph), mag * np.sin(ph))).astype(np.float32)             # :41
    with torch.no_grad():
        est = _n.taylorsenet_forward(sd, torch.from_numpy(feat)[None], cumulative).squeeze(0).numpy()   # :42
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)                  # :44
    eph = np.arctan2(est[1], est[0])
    y = dsp.istft((emag * np.cos(eph) + 1j * emag * np.sin(eph)).T.astype(np.complex64), n_fft, win, hop,
                  length=wav_len)                                           # :48
    taps = {"c": c, "feat": feat, "est": est, "y_norm": y.astype(np.float32)}
    return (y / c).astype(np.float32), taps


def enhance_g2net(sd, wav, p=0.5, cumulative=True):
    """``G2Net_new/com_decode.py:36-88`` (librosa dialect; RECIPROCAL scale convention: c = sqrt(sum x^2 / N), x / c,
    result * c (:43-44,88); p = 0.5 there (:53,76), 1.0 in G2Net_VB/com_decode.py with the InstanceNorm graph):
    compressed RI in, list of stage outputs, the last one decompressed (rule (ii)), iSTFT(length=N)."""
    from . import nets as _n
    n_fft, win, hop = dsp.GEOMETRIES["320"]
    wav = np.asarray(wav, dtype=np.float64)
    c = np.sqrt(np.sum(wav ** 2.0) / len(wav))                              # :43
    x = wav / c                                                             # :44
    spec = dsp.stft(x, n_fft, win, hop).T                                   # :49   [T,161] complex64
    mag, ph = np.abs(spec) ** p, np.angle(spec)                             # :53
    feat = np.stack(((mag * np.cos(ph)).astype(np.float32), (mag * np.sin(ph)).astype(np.float32)))   # :55-56
    with torch.no_grad():
        est = _n.g2net_forward(sd, torch.from_numpy(feat)[None], cumulative)[-1].squeeze(0)   # :68-69  [2,F,T]
    est = est.permute(0, 2, 1).numpy()                                      # :70  [2,T,F]
    emag = np.sqrt(est[0] ** 2 + est[1] ** 2) ** (1.0 / p)                  # :76
    eph = np.arctan2(est[1], est[0])
    de = emag * np.cos(eph) + 1j * emag * np.sin(eph)                       # :77-83
    y = dsp.istft(de.T, n_fft, win, hop, length=len(x))                     # :86-87
    taps = {"c": 1.0 / c, "feat": feat, "est": est, "y_norm": y.astype(np.float32)}   # c in the sqrt(N/sum) convention
    return (y * c).astype(np.float32), taps                                 # :88
