"""The decode loop (wav in -> enhanced wav out) batched on the GPU.

Reference: the ``enhance(args)`` bodies of the ``*_decode*.py`` scripts, e.g.
CRN/crn_decode.py:17-68, LSTM/lstm_decode_vb.py:17-60.  The reference processes one file at a
time and does the DSP on the host; here a batch of equal-length clips stays on the device
from the noisy waveform to the enhanced waveform:

    se_rms_scale -> se_stft (|X|^p and X) -> model.forward -> se_istft (recombine + OLA + 1/c)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from ._lib import ISTFT_CMASK, ISTFT_MAG_PHASE, ISTFT_RI_DECOMP

GEOM_320 = (320, 320, 160)   # LSTM/config.py:4-6, CRN/config.py:4-6


def _lengths_arg(lengths, wav):
    """Per-clip sample counts of a tail-padded batch -> CUDA int32 [B] (None stays None: every clip fills its row).

    Length-aware batching is offered for the families whose network is CAUSAL along time
    (LSTM, CRN, GCRN, DPCRN: uni-directional LSTMs over T, convolutions padded on the past side only, eval BatchNorm,
    per-frame LayerNorm): frame t of the output depends on frames <= t only, so the frames of a clip inside a tail-padded
    batch equal the frames of that clip decoded alone, and only the DSP ends need the clip's own length."""
    if lengths is None:
        return None
    lengths = torch.as_tensor(lengths, dtype=torch.int32).to(wav.device).contiguous()
    if lengths.numel() != wav.shape[0]:
        raise ValueError("one length per clip")
    return lengths


@torch.no_grad()
def enhance_mag_mapping(model, wav, p=1.0, geom=GEOM_320, taps=None, lengths=None):
    """Magnitude-mapping models (LSTM, CRN): backend rule (i) of SURVEY.md section 8(a).
    wav [B,N] float32 CUDA -> enhanced [B,N] float32 CUDA.  CRN/crn_decode.py:38-57.
    ``lengths`` (optional, [B]): clip b has lengths[b] <= N samples and is zero beyond (tail-padded batch); its output
    equals its own decode (samples past its end are 0)."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    n_fft, win, hop = geom
    wav = wav.contiguous().float()
    b, n = wav.shape
    t = 1 + n // hop
    f = n_fft // 2 + 1
    lengths = _lengths_arg(lengths, wav)
    c, inv_c = ops.rms_scale(wav, lengths=lengths)
    mag = torch.empty(b, t, f, device=wav.device, dtype=torch.float32)
    spec = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)   # noisy spectrum (phase)
    ops.stft(wav, c, n_fft, win, hop, mag=mag, re=spec[..., 0], im=spec[..., 1], p_mag=p, lengths=lengths)
    est = model(mag)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_MAG_PHASE, est, None, spec[..., 0], spec[..., 1], n_fft, win, hop, out, n, out_scale=inv_c,
              inv_p=1.0 / p, lengths=lengths)
    if taps is not None:
        taps.update(c=c, mag=mag, spec=spec, est=est)
    return out


def enhance_crn(model, wav, p=1.0, taps=None, lengths=None):
    return enhance_mag_mapping(model, wav, p=p, taps=taps, lengths=lengths)


def enhance_lstm(model, wav, p=1.0, taps=None, lengths=None):
    return enhance_mag_mapping(model, wav, p=p, taps=taps, lengths=lengths)


@torch.no_grad()
def enhance_gcrn(model, wav, p=0.5, taps=None, lengths=None):
    """GCRN/gcrn_decode_vb.py:34-58 (p = 0.5; gcrn_decode.py uses p = 1): compressed real/imag spectrum in,
    real/imag out, |.|^(1/p) with the ESTIMATED phase (backend rule (ii)), iSTFT(length=N), / c.
    wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    n_fft, win, hop = GEOM_320
    wav = wav.contiguous().float()
    b, n = wav.shape
    t, f = 1 + n // hop, n_fft // 2 + 1
    lengths = _lengths_arg(lengths, wav)
    c, inv_c = ops.rms_scale(wav, lengths=lengths)
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)        # compressed RI, channels-last
    ops.stft(wav, c, n_fft, win, hop, re=x[..., 0], im=x[..., 1], p_ri=p, lengths=lengths)
    re, im = model.forward_nhwc(x, taps)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, re, im, None, None, n_fft, win, hop, out, n, out_scale=inv_c, inv_p=1.0 / p,
              lengths=lengths)
    if taps is not None:
        taps.update(c=c, x=x, est=(re, im))
    return out


@torch.no_grad()
def enhance_dpcrn(model, wav, p=1.0, taps=None, lengths=None):
    """DPCRN/dpcrn_decode_vb.py:33-60 (p = 1.0; drcrn_decode.py uses p = 0.5): compressed real/imag spectrum in,
    complex-ratio-masked spectrum out of forward, |.|^(1/p) with its own phase (backend rule (ii)),
    iSTFT(length=N), / c.  wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    n_fft, win, hop = GEOM_320
    wav = wav.contiguous().float()
    b, n = wav.shape
    t, f = 1 + n // hop, n_fft // 2 + 1
    lengths = _lengths_arg(lengths, wav)
    c, inv_c = ops.rms_scale(wav, lengths=lengths)
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)        # compressed RI, channels-last
    ops.stft(wav, c, n_fft, win, hop, re=x[..., 0], im=x[..., 1], p_ri=p, lengths=lengths)
    est = model.forward_nhwc(x, taps)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, est[..., 0], est[..., 1], None, None, n_fft, win, hop, out, n, out_scale=inv_c,
              inv_p=1.0 / p, lengths=lengths)
    if taps is not None:
        taps.update(c=c, x=x, est=est)
    return out


GEOM_FULLSUBNET = (512, 512, 256)   # FullSubNet/fullsubnet_sa_decode.py:53
GEOM_DCCRN = (512, 512, 128)        # DCCRN/dccrn_decode.py:41


@torch.no_grad()
def enhance_dccrn(model, wav, p=0.5, taps=None):
    """DCCRN/dccrn_decode.py:30-60: zero-pad to whole hops, STFT, compress |X|^p with the phase kept,
    DCCRN-E forward (polar mask inside), decompress (rule (ii)), iSTFT without ``length`` then
    ``[:wav_len]``, / c.  wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    n_fft, win, hop = GEOM_DCCRN
    wav = wav.contiguous().float()
    b, n = wav.shape
    c, inv_c = ops.rms_scale(wav)
    frames = -(-n // hop) + 1                       # ceil(N/hop) + 1          (dccrn_decode.py:36)
    fake = (frames - 1) * hop
    if fake != n:                                   # :37-39 zero tail (a no-op when hop divides N)
        padded = torch.zeros(b, fake, device=wav.device, dtype=torch.float32)
        padded[:, :n] = wav
        wav_in = padded
    else:
        wav_in = wav
    t, f = 1 + fake // hop, n_fft // 2 + 1
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)        # compressed RI, channels-last
    ops.stft(wav_in, c, n_fft, win, hop, re=x[..., 0], im=x[..., 1], p_ri=p)
    est = model._forward_nhwc(x, taps)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, est[..., 0], est[..., 1], None, None, n_fft, win, hop, out, n, out_scale=inv_c,
              inv_p=1.0 / p)
    if taps is not None:
        taps.update(c=c, x=x, est=est)
    return out


@torch.no_grad()
def enhance_ctsnet(models, wav, p=1.0, taps=None):
    """CTSNet/two_stage_com_decode_vb.py:61-95 (p = 1.0; 0.5 for the cprs checkpoints): ``models`` =
    (Step1_net, Step2_net).  Zero-pad to whole hops, STFT (compressed RI + magnitude), stage 1 on the magnitude,
    its estimate with the noisy phase, stage 2 on cat(noisy RI, stage-1 RI) + stage-1 RI, decompress (rule (ii)),
    iSTFT without ``length`` then ``[:wav_len]``, / c.  wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    model1, model2 = models
    n_fft, win, hop = GEOM_320
    wav = wav.contiguous().float()
    b, n = wav.shape
    c, inv_c = ops.rms_scale(wav)
    frames = -(-n // hop) + 1                       # ceil(N/hop) + 1   (:65)
    fake = (frames - 1) * hop
    if fake != n:
        padded = torch.zeros(b, fake, device=wav.device, dtype=torch.float32)
        padded[:, :n] = wav
        wav_in = padded
    else:
        wav_in = wav
    t, f = 1 + fake // hop, n_fft // 2 + 1
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)        # compressed RI, channels-last
    mag = torch.empty(b, t, f, device=wav.device, dtype=torch.float32)
    ops.stft(wav_in, c, n_fft, win, hop, mag=mag, re=x[..., 0], im=x[..., 1], p_mag=p, p_ri=p)
    est1 = model1(mag)                                                         # :79
    s2_in = ops.cts_glue1(x, est1)                                             # :80-82
    out_r, out_i = model2.forward_nhwc(s2_in)                                  # :83
    est = ops.cts_glue2(out_r, out_i, s2_in)                                   # :84
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, est[..., 0], est[..., 1], None, None, n_fft, win, hop, out, n, out_scale=inv_c,
              inv_p=1.0 / p)
    if taps is not None:
        taps.update(c=c, x=x, est1=est1, est=est)
    return out


def _padded_to_hops(wav, hop):
    """Zero tail to a whole number of hops (ceil(N/hop) + 1 frames), as the torch-dialect scripts do
    (DCCRN/dccrn_decode.py:36-39, CTSNet/two_stage_com_decode_vb.py:65-68, TaylorSENet/taylorsenet_decode_vb.py:32-35)."""
    b, n = wav.shape
    fake = (-(-n // hop)) * hop
    if fake == n:
        return wav, fake
    padded = torch.zeros(b, fake, device=wav.device, dtype=torch.float32)
    padded[:, :n] = wav
    return padded, fake


@torch.no_grad()
def enhance_taylorsenet(model, wav, p=1.0, taps=None):
    """TaylorSENet/taylorsenet_decode_vb.py:27-52 (p = 1.0; 0.5 for the cprs checkpoints / TaylorSENet_new): STFT
    (compressed RI, channels-last), forward (RI rows out), decompress (rule (ii)), iSTFT(length=N), / c."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    from .taylor import N_BINS, RI_LD
    n_fft, win, hop = GEOM_320
    wav = wav.contiguous().float()
    b, n = wav.shape
    c, inv_c = ops.rms_scale(wav)
    wav_in, fake = _padded_to_hops(wav, hop)
    t, f = 1 + fake // hop, n_fft // 2 + 1
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)
    ops.stft(wav_in, c, n_fft, win, hop, re=x[..., 0], im=x[..., 1], p_ri=p)
    rows = model.forward_nhwc(x, taps).view(b, t, RI_LD)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, rows[..., :N_BINS], rows[..., N_BINS:2 * N_BINS], None, None, n_fft, win, hop, out, n,
              out_scale=inv_c, inv_p=1.0 / p)
    if taps is not None:
        taps.update(c=c, x=x, est=rows)
    return out


@torch.no_grad()
def enhance_g2net(model, wav, p=0.5, taps=None):
    """G2Net_new/com_decode.py:36-88 (p = 0.5; 1.0 for G2Net_VB): x / sqrt(sum x^2 / N) (the reciprocal spelling of the
    same RMS scale), STFT (compressed RI), gaf_base forward, last stage decompressed (rule (ii)), iSTFT(length=N), * c.
    wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    from .g2net import IM_OFF, N_BINS, RI_LD
    n_fft, win, hop = GEOM_320
    wav = wav.contiguous().float()
    b, n = wav.shape
    t, f = 1 + n // hop, n_fft // 2 + 1
    c, inv_c = ops.rms_scale(wav)                    # c = sqrt(N / sum x^2) = 1 / (the script's c); inv_c = the script's c
    x = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)
    ops.stft(wav, c, n_fft, win, hop, re=x[..., 0], im=x[..., 1], p_ri=p)
    rows = model.forward_nhwc(x, taps).view(b, t, RI_LD)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_RI_DECOMP, rows[..., :N_BINS], rows[..., IM_OFF:IM_OFF + N_BINS], None, None, n_fft, win, hop, out, n,
              out_scale=inv_c, inv_p=1.0 / p)
    if taps is not None:
        taps.update(c=c, x=x, est=rows)
    return out


@torch.no_grad()
def enhance_fullsubnet(model, wav, p=0.5, taps=None):
    """FullSubNet/fullsubnet_sa_decode.py:44-78: |X|^p magnitude in, complex mask out, mask applied
    to the COMPRESSED spectrum, decompressed, iSTFT(length=N), / c  (backend rule (iii)).
    wav [B,N] float32 CUDA -> [B,N]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    n_fft, win, hop = GEOM_FULLSUBNET
    wav = wav.contiguous().float()
    b, n = wav.shape
    t, f = 1 + n // hop, n_fft // 2 + 1
    c, inv_c = ops.rms_scale(wav)
    mag = torch.empty(b, 1, f, t, device=wav.device, dtype=torch.float32)     # [B,1,F,T] as the script feeds
    spec = torch.empty(b, 2, f, t, device=wav.device, dtype=torch.float32)    # uncompressed X
    ops.stft(wav, c, n_fft, win, hop, mag=mag[:, 0], re=spec[:, 0], im=spec[:, 1], layout="bft", p_mag=p)
    mask = model(mag)                                                          # [B,2,F,T] (strided view)
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_CMASK, mask[:, 0], mask[:, 1], spec[:, 0], spec[:, 1], n_fft, win, hop, out, n,
              out_scale=inv_c, inv_p=1.0 / p, p_x=p, layout_a="bft", layout_b="bft")
    if taps is not None:
        taps.update(c=c, mag=mag, mask=mask)
    return out


@torch.no_grad()
def enhance_uformer(model, wav, taps=None):
    """Uformer/uformer_decode.py:38-50: c-normalise, model(x, x) (STFT 512/400/160 and iSTFT inside the
    model, output length hop*(T-1)), / c.  wav [B,N] -> [B, hop*(N//hop)]."""
    if not wav.is_cuda:
        raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
    wav = wav.contiguous().float()
    c, inv_c = ops.rms_scale(wav)
    out, _, est, _ = model(wav, None, scale=c, out_scale=inv_c)
    if taps is not None:
        taps.update(c=c, est=est)
    return out


@torch.no_grad()
def dsp_roundtrip(wav, geom):
    """STFT -> identity -> iSTFT (the DSP-only run of SURVEY.md section 8(d))."""
    from ._lib import ISTFT_SPEC
    n_fft, win, hop = geom
    b, n = wav.shape
    t = 1 + n // hop
    f = n_fft // 2 + 1
    c, inv_c = ops.rms_scale(wav)
    spec = torch.empty(b, t, f, 2, device=wav.device, dtype=torch.float32)
    ops.stft(wav, c, n_fft, win, hop, mag=None, re=spec[..., 0], im=spec[..., 1])
    out = torch.empty(b, n, device=wav.device, dtype=torch.float32)
    ops.istft(ISTFT_SPEC, spec[..., 0], spec[..., 1], None, None, n_fft, win, hop, out, n, out_scale=inv_c)
    return out


# ---------------------------------------------------------------------------------------------
# File-level contract of the decode scripts: every file of a directory in, same basename out.
# soundfile is not available here; 16-bit PCM wav I/O through scipy (sf.write's default
# subtype for .wav is PCM_16 too, CRN/crn_decode.py:67).
# ---------------------------------------------------------------------------------------------
def enhancer_for(model):
    """The decode loop that goes with a drop-in model instance (or the (Step1_net, Step2_net) pair of CTSNet)."""
    import sys
    pkg = sys.modules[__package__]          # class names as the package exports them (dpcrn is the class, as in DPCRN.py)
    if isinstance(model, (tuple, list)):
        return enhance_ctsnet
    table = [(pkg.crn_net, enhance_crn), (pkg.lstm_net, enhance_lstm), (pkg.gcrn.Net, enhance_gcrn),
             (pkg.dpcrn, enhance_dpcrn), (pkg.DCCRN, enhance_dccrn), (pkg.fullsubnet.Model, enhance_fullsubnet),
             (pkg.Uformer, enhance_uformer), (pkg.TaylorSENet, enhance_taylorsenet),
             (pkg.g2net.gaf_base, enhance_g2net)]
    for cls, fn in table:
        if isinstance(model, cls):
            return fn
    raise TypeError(f"no decode loop registered for {type(model).__name__}")


class GraphedEnhance:
    """A decode loop captured once per input shape in a CUDA graph and replayed: one ``cudaGraphLaunch`` instead of the
    350 .. 1 900 kernel launches (26 .. 76 us of Python + ctypes each) a batch of the TCM-family / FullSubNet models
    costs -- for those the GPU otherwise waits for the host (tools/host_bound_check.py).  The reference has no
    counterpart: it re-dispatches every ATen op per utterance (CRN/crn_decode.py:37-57).

        dec = GraphedEnhance(model)                 # or GraphedEnhance((step1, step2)) for CTSNet; kw -> decode loop
        y = dec(wav)                                # [B,N] float32 CUDA; y is valid until the next call with this shape

    Shapes are static inside a graph: every distinct (B, N, ragged?) gets its own capture (a warm-up call, then the
    capture, both on a side stream); ``lengths`` (per-clip sample counts, time-causal families) are copied into a
    static tensor, so ragged batches of one padded shape share a graph."""

    def __init__(self, model, enhance_fn=None, **kw):
        self.model, self.fn, self.kw = model, enhance_fn or enhancer_for(model), kw
        self._graphs = {}

    def _capture(self, wav, lengths):
        static_in = wav.clone()
        static_len = None if lengths is None else lengths.clone()
        kw = dict(self.kw)
        if static_len is not None:
            kw["lengths"] = static_len
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.fn(self.model, static_in, **kw)          # packs weights, fills lazy tables, sizes the allocator pools
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self.fn(self.model, static_in, **kw)
        return graph, static_in, static_len, static_out

    @torch.no_grad()
    def __call__(self, wav, lengths=None):
        if not wav.is_cuda:
            raise RuntimeError("se_b200.decode needs CUDA tensors (no CPU fallback)")
        wav = wav.contiguous().float()
        if lengths is not None:
            lengths = _lengths_arg(lengths, wav)
        key = (tuple(wav.shape), wav.device, lengths is not None)
        if key not in self._graphs:
            self._graphs[key] = self._capture(wav, lengths)
        graph, static_in, static_len, static_out = self._graphs[key]
        static_in.copy_(wav)
        if static_len is not None:
            static_len.copy_(lengths)
        graph.replay()
        return static_out


def enhance_host_stream(model, host_batches, enhance_fn=None, depth=2, device=None, **kw):
    """Pipelined host -> host decode: for every pinned host batch [B,N] float32 of ``host_batches`` yields the enhanced
    batch as a pinned host tensor that stays valid until ``depth`` MORE batches have been drawn from the generator
    (ring of ``2 * depth`` pinned output buffers: the download of batch i + depth is queued before batch i + 1 is yielded,
    so a ring of ``depth`` would be overwritten one draw later).  The upload of batch i+1 and the download of batch i-1
    run on their own CUDA streams while batch i is in the decode loop, so the PCIe copies the reference pays serially
    per utterance (CRN/crn_decode.py:46-47,53) disappear behind the compute.  All batches must have the same shape."""
    fn = enhance_fn or enhancer_for(model)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    compute = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    nout = 2 * depth
    dev_in, host_out = [None] * depth, [None] * nout
    ev_in = [torch.cuda.Event() for _ in range(depth)]
    ev_done = [None] * depth
    ev_out = [None] * nout
    pending = []                                    # output slots whose download has been queued, oldest first
    for i, host in enumerate(host_batches):
        slot, oslot = i % depth, i % nout
        if len(pending) == depth:                   # the consumer gets batch i - depth before more work is queued
            o = pending.pop(0)
            ev_out[o].synchronize()
            yield host_out[o]
        if dev_in[slot] is None:
            dev_in[slot] = torch.empty(host.shape, device=dev, dtype=torch.float32)
        if host_out[oslot] is None:
            # straight from the pinned caching allocator: .pin_memory() on a pageable tensor first touches and copies 16 MB
            # per buffer (about 0.5 ms per step of a 30-step run went there)
            host_out[oslot] = torch.empty(host.shape, dtype=torch.float32, pin_memory=True)
        with torch.cuda.stream(s_in):
            if ev_done[slot] is not None:
                s_in.wait_event(ev_done[slot])      # the decode loop of batch i - depth has finished reading this buffer
            dev_in[slot].copy_(host, non_blocking=True)
            ev_in[slot].record(s_in)
        compute.wait_event(ev_in[slot])
        y = fn(model, dev_in[slot], **kw)
        ev_done[slot] = torch.cuda.Event()
        ev_done[slot].record(compute)
        y.record_stream(s_out)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[slot])
            host_out[oslot].copy_(y, non_blocking=True)   # last handed out 2 * depth batches ago: depth draws have passed
            ev_out[oslot] = torch.cuda.Event()
            ev_out[oslot].record(s_out)
        pending.append(oslot)
    for o in pending:
        ev_out[o].synchronize()
        yield host_out[o]


def read_wav_any(path):
    """``soundfile.read`` as the scripts use it (CRN/crn_decode.py:38): (float64 samples in [-1, 1), sample rate)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.ndim != 1:
        raise ValueError(f"{path}: {x.ndim}-D audio; the decode scripts handle mono files only")
    if x.dtype.kind == "i":
        x = x.astype(np.float64) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":                                      # 8-bit PCM is unsigned
        x = (x.astype(np.float64) - 128.0) / 128.0
    return x.astype(np.float64), int(sr)


def read_wav(path, fs):
    """read_wav_any for the scripts that do not resample (CRN/crn_decode.py:38): the file must already be at ``fs``."""
    x, sr = read_wav_any(path)
    if sr != fs:
        raise ValueError(f"{path}: sample rate {sr} != {fs} (pass resample_to= for the librosa.resample front step of "
                         "the *_decode_vb.py scripts, lstm_decode_vb.py:34)")
    return x


def write_wav(path, y, fs):
    """``soundfile.write(path, y, fs)`` with its default subtype for .wav: 16-bit PCM (CRN/crn_decode.py:67).
    libsndfile converts float -> short as ``lrint(x * 0x7FFF)`` (and short -> float as ``x / 0x8000`` on read)."""
    from scipy.io import wavfile
    pcm = np.clip(np.rint(np.asarray(y, dtype=np.float64) * 32767.0), -32768, 32767).astype(np.int16)
    wavfile.write(path, fs, pcm)


def _wav_header(path):
    """(sample rate, samples) of a mono wav file without reading its audio (memory-mapped)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path, mmap=True)
    if x.ndim != 1:
        raise ValueError(f"{x.ndim}-D audio; the decode scripts handle mono files only")
    return int(sr), int(x.shape[0])


def plan_batches(infos, batch, ragged, pad_tolerance=0.25, min_len=512):
    """Deterministic batch list for ``enhance_dir``.  infos: [(name, sample rate, samples)].  ``ragged`` (the decode loop
    takes per-clip ``lengths``): files of one sample rate are sorted by length and cut into batches of at most ``batch``
    files whose longest member is at most ``1 + pad_tolerance`` times the shortest (bounded padding waste); otherwise
    only files of identical (rate, length) share a batch.  Returns [(sr, [names], [lengths])]."""
    out = []
    if not ragged:
        groups = {}
        for name, sr, n in infos:
            groups.setdefault((sr, n), []).append(name)
        for (sr, n), names in sorted(groups.items()):
            for i in range(0, len(names), batch):
                out.append((sr, names[i:i + batch], [n] * len(names[i:i + batch])))
        return out
    by_sr = {}
    for name, sr, n in infos:
        by_sr.setdefault(sr, []).append((n, name))
    for sr, items in sorted(by_sr.items()):
        items.sort()
        cur = []
        for n, name in items:
            if cur and (len(cur) >= batch or n > (1.0 + pad_tolerance) * cur[0][0] or cur[0][0] < min_len):
                out.append((sr, [x[1] for x in cur], [x[0] for x in cur]))
                cur = []
            cur.append((n, name))
        if cur:
            out.append((sr, [x[1] for x in cur], [x[0] for x in cur]))
    return out


def enhance_dir(model, mix_file_path, esti_file_path, fs=16000, batch=64, device="cuda", enhance_fn=None,
                resample_to=None, rank=None, world=None, pad_tolerance=0.25, report=None, **kw):
    """wav directory in -> wav directory out: the ``enhance(args)`` surface of the decode scripts (``args.mix_file_path``,
    ``args.esti_clean_file_path`` / ``args.esti_file_path``, ``args.fs``; e.g. CRN/crn_decode_vb.py:17-64).  The
    reference loops one file at a time, any length (crn_decode_vb.py:31-33); here files are batched and a batch stays on
    the device from the noisy waveform to the enhanced one:

    * decode loops that take per-clip ``lengths`` (the time-causal families LSTM, CRN, GCRN, DPCRN) get LENGTH-BUCKETED
      batches: files sorted by length, tail-padded to the longest of their batch (at most ``pad_tolerance`` longer than
      the shortest), every clip decoded exactly as if alone (``plan_batches``);
    * the other families (utterance-level statistics or attention over time, look-ahead) batch files of identical length;
    * ``resample_to=16000`` adds the front step of the ``*_decode_vb.py`` scripts (``librosa.resample(x, orig_fs, 16000,
      fix=True, scale=False)``, LSTM/lstm_decode_vb.py:33-34) on the device; those files are grouped by (rate, length).

    Only wav headers are read up front; audio is loaded per batch, for this rank's batches only, one batch ahead on a
    reader thread.  Entries that are not readable mono wav files are skipped and reported, never fatal (a script
    looping over os.listdir would die on the first one).  ``rank`` / ``world`` (default: the
    initialised ``torch.distributed`` group, else one rank) shard the batch list over processes, one per GPU: every rank
    builds the same list and decodes batches ``rank, rank + world, ...``; utterances are independent and every rank
    writes its own files, so no collective is involved.  Returns the number of files THIS rank
    wrote; ``report`` (a dict) receives the details.  ``kw`` goes to the decode loop (``p=0.5`` for the compressed
    checkpoints)."""
    import inspect
    from concurrent.futures import ThreadPoolExecutor
    if world is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    fn = enhance_fn or enhancer_for(model)
    ragged = resample_to is None and "lengths" in inspect.signature(fn).parameters
    os.makedirs(esti_file_path, exist_ok=True)
    infos, skipped = [], {}
    for name in sorted(os.listdir(mix_file_path)):
        path = os.path.join(mix_file_path, name)
        if not os.path.isfile(path):
            skipped[name] = "not a file"
            continue
        try:
            sr, n = _wav_header(path)
        except Exception as e:                       # not a wav file / unsupported encoding / truncated header
            skipped[name] = f"unreadable: {e}"
            continue
        if resample_to is None and sr != fs:
            skipped[name] = f"sample rate {sr} != {fs} (pass resample_to=)"
            continue
        infos.append((name, sr, n))
    batches = plan_batches(infos, batch, ragged, pad_tolerance)
    mine = [bt for i, bt in enumerate(batches) if i % world == rank]

    def load(bt):
        sr, names, lens = bt
        nmax = max(lens)
        buf = np.zeros((len(names), nmax), dtype=np.float32)
        ok = []
        for i, (name, n) in enumerate(zip(names, lens)):
            try:
                x, _ = read_wav_any(os.path.join(mix_file_path, name))
                if len(x) != n:
                    raise ValueError(f"{len(x)} samples, header says {n}")
                buf[i, :n] = x
                ok.append(i)
            except Exception as e:                   # a file that turns out to be corrupt does not poison its batch
                skipped[name] = f"unreadable: {e}"
        return buf, ok

    count, padded, total = 0, 0, 0
    written = []
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(load, mine[0]) if mine else None
        for k, (sr, names, lens) in enumerate(mine):
            buf, ok = fut.result()
            fut = pool.submit(load, mine[k + 1]) if k + 1 < len(mine) else None
            if not ok:
                continue
            if len(ok) != len(names):
                buf = buf[ok]
                names, lens = [names[i] for i in ok], [lens[i] for i in ok]
            wav = torch.from_numpy(buf).to(device)
            if resample_to is not None:
                wav = ops.resample(wav, sr, resample_to)
            if ragged and len(set(lens)) > 1:
                out = fn(model, wav, lengths=torch.tensor(lens, dtype=torch.int32), **kw)
            else:
                out = fn(model, wav, **kw)
            out = out.cpu().numpy()
            for name, n, y in zip(names, lens, out):
                write_wav(os.path.join(esti_file_path, name), y if resample_to is not None else y[:n], fs)
                written.append(name)
                count += 1
            padded += len(names) * max(lens) - sum(lens)
            total += len(names) * max(lens)
    if report is not None:
        report.update(written=written, skipped=skipped, batches=len(mine), batches_total=len(batches), ragged=ragged,
                      padded_fraction=(padded / total if total else 0.0))
    return count


def enhance(args, model, **kw):
    """Drop-in for the scripts' ``enhance(args)`` given an already constructed / loaded model (the scripts hard-code the
    checkpoint path, e.g. CRN/crn_decode.py:19): reads ``args.mix_file_path``, writes ``args.esti_clean_file_path`` (or
    ``args.esti_file_path``, as FullSubNet / CTSNet / G2Net / TaylorSENet spell it) at ``args.fs``."""
    out_dir = getattr(args, "esti_clean_file_path", None) or getattr(args, "esti_file_path")
    n = enhance_dir(model, args.mix_file_path, out_dir, fs=getattr(args, "fs", 16000), **kw)
    for i in range(n):
        print(' The %d utterance has been decoded!' % (i + 1))
    return n
