"""Stateful streaming decode for the time-causal magnitude-mapping models.

The reference decodes whole files (CRN/crn_decode.py:37-57); its CRN and LSTM networks are causal (uni-directional
LSTMs CRN/CRN.py:20,29, LSTM/LSTM.py:17-18; convolutions padded on the past side only, CRN.py:38,112-117), so the same
output can be produced chunk by chunk as the audio arrives.  ``MagStream`` does that for B parallel streams and is
EXACT: the concatenation of what ``push`` / ``flush`` return equals the offline ``decode.enhance_crn`` /
``enhance_lstm`` of the whole clips (tests/test_gpu_models.py::test_streaming_equals_offline), given the same RMS
constant ``c`` -- the one quantity of the decode loop that is not causal (crn_decode.py:39 needs the whole file; a live
caller supplies a running / calibrated level instead).

State carried between chunks, all on the device:
  * STFT: the input samples that later frames still need (se_stft is run on [carried samples | new samples] as a local
    clip and only frames whose window lies inside are kept; the true start / end use its reflect padding);
  * network: (h, c) of every LSTM layer, advanced one frame per se_lstm_cell_tf32x3_ex launch (the "lstm step" entry of
    the C ABI); for CRN also the last 10 magnitude frames and the last 5 LSTM output frames, from which the encoder /
    decoder convolutions (one frame of past per layer, 5 + 5 layers) recompute their context -- no kernel needs a
    streaming variant;
  * iSTFT: the last n_fft / hop - 1 estimated / noisy frames, so that every emitted sample has all its overlapping
    frames and the same window envelope as offline.
Latency = one chunk + n_fft / 2 samples of look-ahead (centre = True framing).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import ISTFT_MAG_PHASE


class _LstmState:
    """(h_hi, h_lo, c) of one layer for M streams, h double-buffered (the cell kernel must not overwrite its input)."""

    def __init__(self, m, hidden, device):
        z = lambda: torch.zeros(m, hidden, device=device, dtype=torch.float32)   # noqa: E731
        self.h = [(z(), z()), (z(), z())]
        self.c = z()
        self.cur = 0
        self.first = True

    def step(self, x_pair, cell, h_out=None):
        src, dst = self.h[self.cur], self.h[self.cur ^ 1]
        ops.lstm_cell_tf32x3_ex(x_pair, None if self.first else src, cell["w_hi"], cell["w_lo"], cell["bias"], self.c,
                                dst[0], dst[1], h_out)
        self.cur ^= 1
        self.first = False
        return dst


class MagStream:
    """Streaming decode of B parallel streams through a ``crn_net`` or ``lstm_net`` (320 / 320 / 160 geometry).

        st = MagStream(model, c)                    # c [B] float32 CUDA: the RMS constants (decode.ops.rms_scale)
        y0 = st.push(x[:, :n0]); y1 = st.push(x[:, n0:n1]); ...; yl = st.flush()
        torch.cat([y0, y1, ..., yl], 1) == decode.enhance_crn(model, x)          (to fp32 rounding)

    ``push`` takes [B, n] float32 CUDA samples (any n >= 0) and returns the enhanced samples that became final."""

    def __init__(self, model, c, inv_c=None, p=1.0, geom=(320, 320, 160)):
        from .crn import crn_net
        from .lstm import lstm_net
        if not isinstance(model, (crn_net, lstm_net)):
            raise TypeError("MagStream covers the time-causal magnitude-mapping models (crn_net, lstm_net)")
        self.model, self.is_crn = model, isinstance(model, crn_net)
        self.n_fft, self.win, self.hop = geom
        self.h2 = self.n_fft // 2
        self.p = float(p)
        self.c = c.contiguous().float()
        self.inv_c = (1.0 / self.c) if inv_c is None else inv_c.contiguous().float()
        self.b = self.c.numel()
        dev = self.c.device
        model._ensure_packed()
        self.ctx_enc, self.ctx_dec = model.STREAM_CONTEXT if self.is_crn else (0, 0)
        self.cells = model.stream_cells()
        self.states = [_LstmState(self.b, 1024, dev) for _ in self.cells]
        self.f = self.n_fft // 2 + 1
        self.buf = torch.empty(self.b, 0, device=dev)      # samples [buf_start, received)
        self.buf_start = 0
        self.received = 0
        self.frames_done = 0                               # global frames already through the network
        self.mag_hist = torch.empty(self.b, 0, self.f, device=dev)        # last ctx_enc + ctx_dec magnitude frames
        self.lstm_hist = torch.empty(self.b, 0, 1024, device=dev)         # last ctx_dec LSTM output frames
        self.kf = self.n_fft // self.hop                   # frames overlapping one sample
        self.est_hist = torch.empty(self.b, 0, self.f, device=dev)        # last kf - 1 estimated magnitude frames
        self.spec_hist = torch.empty(self.b, 0, self.f, 2, device=dev)    # ... and their noisy spectra
        self.samples_out = 0

    # ---- network on a run of new frames -------------------------------------------------------------
    def _lstm_run(self, x_pair_of_frame, nframes):
        """Advance all LSTM layers over ``nframes`` frames; returns the last layer's h as [B, nframes, 1024]."""
        out = torch.empty(nframes, self.b, 1024, device=self.c.device)      # frame-major: out[i] has the state's row stride
        for i in range(nframes):
            pair = x_pair_of_frame(i)
            for l, (cell, st) in enumerate(zip(self.cells, self.states)):
                last = l == len(self.cells) - 1
                pair = st.step(pair, cell, out[i] if last else None)
        return out.transpose(0, 1).contiguous()

    def _network(self, mag_new):
        b, tc, _ = mag_new.shape
        dev = mag_new.device
        if not self.is_crn:
            return self._network_lstm(mag_new)
        model = self.model
        k = self.mag_hist.shape[1]
        mag = torch.cat([self.mag_hist, mag_new], 1).contiguous()           # local frames [0, k + tc)
        enc = model._encoder(mag, f16=False)      # the streaming path stays on TF32 pairs (its cells are packed that way)
        e5 = enc[-1]
        if e5.pair is None:
            e5.pair = ops.split_tf32(e5.f32)
        hi, lo = (x.view(b, k + tc, 1024) for x in e5.pair)
        hs_new = self._lstm_run(lambda i: (hi[:, k + i], lo[:, k + i]), tc)
        # decoder over all local frames; LSTM outputs older than the carried ones only feed discarded frames
        kl = self.lstm_hist.shape[1]
        hs = torch.zeros(b, k + tc, 1024, device=dev)
        hs[:, k - kl:k] = self.lstm_hist
        hs[:, k:] = hs_new
        est = model._decoder(hs, enc, f16=False)[:, k:]
        keep = self.ctx_enc + self.ctx_dec
        self.mag_hist = mag[:, max(0, k + tc - keep):].contiguous()
        self.lstm_hist = hs[:, max(0, k + tc - self.ctx_dec):].contiguous()
        return est

    def _network_lstm(self, mag_new):
        """lstm_net (LSTM/LSTM.py:24-29): BatchNorm folded into the first projection, three LSTM layers, Linear + Softplus."""
        b, tc, f = mag_new.shape
        P = self.model._packed
        kx = self.cells[0]["kx"]
        hi, lo = ops.pad_split_tf32(mag_new.reshape(b * tc, f).contiguous(), kx)
        hi, lo = hi.view(b, tc, kx), lo.view(b, tc, kx)
        hs = self._lstm_run(lambda i: (hi[:, i], lo[:, i]), tc)
        if b * tc >= 128:                                        # same engines as lstm_net._forward_impl
            a_hi, a_lo = ops.split_tf32(hs.view(b * tc, 1024))
            y = ops.gemm_tf32x3(a_hi, a_lo, P["fc_hi"], P["fc_lo"], P["fc_b"], 161, act="softplus")
        else:
            y = ops.linear(hs.view(b * tc, 1024), P["fc_w"], P["fc_b"], 161, act="softplus")
        return y.view(b, tc, 161)

    # ---- DSP ends --------------------------------------------------------------------------------------
    def _process_frames(self, t1, final_len=None):
        """Bring global frames [frames_done, t1] through STFT -> network -> iSTFT; returns the newly final samples."""
        t0 = self.frames_done
        if t1 < t0:
            return torch.empty(self.b, 0, device=self.c.device)
        hop, h2 = self.hop, self.h2
        m = -(-h2 // hop)                                      # frames of left margin whose window may touch the local start
        a = max(0, (t0 - m) * hop)                             # local clip = global samples [a, e)
        e = self.received if final_len is not None else t1 * hop + h2
        if e - a < self.n_fft:                                 # se_stft wants a whole window: widen (extra frames are dropped)
            e = min(self.received, a + self.n_fft)
            a = max(0, min(a, ((e - self.n_fft) // hop) * hop))
        clip = self.buf[:, a - self.buf_start:e - self.buf_start].contiguous()
        if clip.shape[1] < self.n_fft:
            raise ValueError("a stream needs at least n_fft samples before its first frames can be produced")
        tl = 1 + clip.shape[1] // hop
        mag = torch.empty(self.b, tl, self.f, device=clip.device)
        spec = torch.empty(self.b, tl, self.f, 2, device=clip.device)
        ops.stft(clip, self.c, self.n_fft, self.win, hop, mag=mag, re=spec[..., 0], im=spec[..., 1], p_mag=self.p)
        j0 = t0 - a // hop
        ntc = t1 - t0 + 1
        mag_new, spec_new = mag[:, j0:j0 + ntc].contiguous(), spec[:, j0:j0 + ntc]
        est_new = self._network(mag_new)
        # iSTFT on [carried kf - 1 frames | new frames] as a local clip starting at global frame ta
        kh = self.est_hist.shape[1]
        ta = t0 - kh
        est = torch.cat([self.est_hist, est_new], 1).contiguous()
        sp = torch.cat([self.spec_hist, spec_new], 1).contiguous()
        if final_len is None:
            lloc = (t1 + 1 - ta) * hop - h2                    # local samples whose covering frames are all present
        else:
            lloc = final_len - ta * hop
        out = torch.empty(self.b, max(lloc, 0), device=clip.device)
        if lloc > 0:
            ops.istft(ISTFT_MAG_PHASE, est, None, sp[..., 0], sp[..., 1], self.n_fft, self.win, hop, out, lloc,
                      out_scale=self.inv_c, inv_p=1.0 / self.p)
        first_new = self.samples_out - ta * hop                # local index of the first sample not yet emitted
        y = out[:, max(first_new, 0):]
        self.samples_out += y.shape[1]
        keep = self.kf - 1
        self.est_hist = est[:, max(0, est.shape[1] - keep):].contiguous()
        self.spec_hist = sp[:, max(0, sp.shape[1] - keep):].contiguous()
        self.frames_done = t1 + 1
        # samples before the earliest possible start of the next local clip are no longer needed
        na = max(0, (self.frames_done - m) * hop - self.n_fft)
        if na > self.buf_start:
            self.buf = self.buf[:, na - self.buf_start:].contiguous()
            self.buf_start = na
        return y

    @torch.no_grad()
    def push(self, samples):
        if samples.shape[0] != self.b or not samples.is_cuda:
            raise ValueError("push expects [B, n] CUDA samples for the B streams this object was created with")
        self.buf = torch.cat([self.buf, samples.float()], 1)
        self.received += samples.shape[1]
        t1 = (self.received - self.h2) // self.hop if self.received >= self.h2 else -1   # frames with a complete window
        if self.received < self.n_fft or t1 < self.frames_done:
            return torch.empty(self.b, 0, device=samples.device)
        return self._process_frames(t1)

    @torch.no_grad()
    def flush(self):
        """End of the streams: the remaining frames (reflect padding at the true end) and the tail of the output."""
        t_last = self.received // self.hop                     # T - 1 with T = 1 + N // hop
        return self._process_frames(t_last, final_len=self.received)
