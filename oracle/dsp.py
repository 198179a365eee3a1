"""Oracle DSP: STFT / iSTFT / RMS-scale exactly as the reference decode scripts use them.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy float64 throughout,
rounded where the reference rounds.

Two dialects exist in the reference (layer L2):

* librosa dialect -- ``librosa.stft(y, n_fft, hop_length, window='hanning')`` on a
  float64 waveform, result cast to complex64; ``librosa.istft(S, hop_length,
  win_length, window='hanning', length=N)`` with float32 output
  (``CRN/crn_decode.py:41,55-56``, ``LSTM/lstm_decode_vb.py:37,50-51``,
  ``GCRN/gcrn_decode.py:42,61-62``, ``DPCRN/drcrn_decode.py:42,61-62``).
  librosa is un-vendored and un-pinned; the algorithm restated here is the
  published librosa<=0.7 one: centre reflect-pad of n_fft/2, periodic Hann padded
  (centred) to n_fft, one-sided FFT; inverse = irFFT * window, overlap-add,
  divide by the window-sum-square envelope where it exceeds ``tiny``, drop the
  first n_fft/2 samples, fix the length.
* torch dialect -- ``torch.stft(x, n_fft, hop, win, hann_window)`` /
  ``torch.istft`` on a float32 waveform (``DCCRN/dccrn_decode.py:41,56``,
  ``FullSubNet/fullsubnet_sa_decode.py:53,76``, ``Uformer/uformer.py:178,276``).
  Same mathematics; the waveform is rounded to float32 first and the spectrum is
  complex64.

Parity status: no reference test pins these numerics ("parity unpinned" by the
reference itself); ``tests/test_oracle_dsp.py`` pins this restatement against
``torch.stft``/``torch.istft`` for all four geometries.
"""
from __future__ import annotations

import numpy as np

GEOMETRIES = {
    # name: (n_fft, win, hop)      -- SURVEY.md section 0.1 / 8(a) descriptor table
    "320": (320, 320, 160),        # LSTM, CRN, GCRN, DPCRN, CTSNet, G2Net, TaylorSENet
    "dccrn": (512, 512, 128),      # DCCRN/dccrn_decode.py:41
    "fullsubnet": (512, 512, 256), # FullSubNet/fullsubnet_sa_decode.py:53
    "uformer": (512, 400, 160),    # Uformer/uformer.py:33-35
}


def hann_periodic(win: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win, fftbins=True) == torch.hann_window(win)."""
    n = np.arange(win, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win)


def padded_window(n_fft: int, win: int) -> np.ndarray:
    """Window zero-padded, centred, to n_fft (librosa.util.pad_center; torch.stft does the same)."""
    w = np.zeros(n_fft, dtype=np.float64)
    left = (n_fft - win) // 2
    w[left:left + win] = hann_periodic(win)
    return w


def num_frames(n: int, hop: int) -> int:
    """centre=True framing: 1 + N // hop."""
    return 1 + n // hop


def rms_scale(wav: np.ndarray):
    """c = sqrt(N / sum x^2); x*c   (CRN/crn_decode.py:39-40). float64."""
    wav = np.asarray(wav, dtype=np.float64)
    c = np.sqrt(len(wav) / np.sum(wav ** 2.0))
    return wav * c, c


def stft(y: np.ndarray, n_fft: int, win: int, hop: int, out_dtype=np.complex64) -> np.ndarray:
    """Returns [F, T] like librosa.stft / torch.stft(return_complex=True)."""
    y = np.asarray(y, dtype=np.float64)
    w = padded_window(n_fft, win)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    t = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(t)[:, None]
    frames = yp[idx] * w[None, :]
    spec = np.fft.rfft(frames, axis=1)            # [T, F]
    return spec.T.astype(out_dtype)


def istft(spec: np.ndarray, n_fft: int, win: int, hop: int, length: int | None,
          out_dtype=np.float32) -> np.ndarray:
    """spec [F, T] -> waveform.  length=None trims n_fft/2 on both sides (torch / librosa default)."""
    spec = np.asarray(spec)
    t = spec.shape[1]
    w = padded_window(n_fft, win)
    frames = np.fft.irfft(spec.T.astype(np.complex128), n=n_fft, axis=1) * w[None, :]
    total = n_fft + hop * (t - 1)
    y = np.zeros(total, dtype=np.float64)
    env = np.zeros(total, dtype=np.float64)
    w2 = w * w
    for i in range(t):
        y[i * hop:i * hop + n_fft] += frames[i]
        env[i * hop:i * hop + n_fft] += w2
    nz = env > np.finfo(np.float32).tiny
    y[nz] /= env[nz]
    start = n_fft // 2
    if length is None:
        y = y[start:total - start]
    else:
        y = y[start:]
        if len(y) >= length:
            y = y[:length]
        else:
            y = np.pad(y, (0, length - len(y)))
    return y.astype(out_dtype)
