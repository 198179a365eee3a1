"""Oracle for the 48 kHz -> 16 kHz front step of the ``*_decode_vb.py`` scripts.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Reference call sites: ``LSTM/lstm_decode_vb.py:33-34``, ``DCCRN/dccrn_decode_vb.py:25-26``::

    feat_wav, orig_fs = sf.read(path)
    feat_wav = librosa.resample(feat_wav, orig_fs, 16000, fix=True, scale=False)

**Parity unpinned by the reference; pinned here against an independent implementation.**  ``librosa`` (positional ``resample(y, orig_sr, target_sr, fix=, scale=)`` => librosa <= 0.7) and
its back end ``resampy`` (``res_type='kaiser_best'`` is librosa's default) are third-party dependencies that are neither
vendored in ``/root/reference`` nor installed here, with no version pin and no test or golden vector in the reference.
What follows restates their PUBLISHED algorithm (resampy 0.2.x, J. O. Smith's band-limited interpolation):

* ``sinc_window``   -- resampy/filters.py: right wing of ``rolloff * sinc(rolloff * t)`` on a grid of ``2**precision``
  samples per zero crossing, tapered by the right half of a symmetric Kaiser window;
* ``kaiser_best``   -- the parameters resampy documents for its shipped ``kaiser_best`` table: 64 zero crossings,
  precision 9, rolloff 0.9475937167399596, beta 14.769656459379492 (the shipped ``.npz`` is that function's output);
* ``resample_f``    -- resampy/interpn.py: per output sample, left and right filter wings with linear interpolation
  between table entries (``interp_delta``), a float64 ``time_register`` advanced by ``1 / ratio``;
* ``resampy_resample`` -- resampy/core.py: output length ``int(n * ratio)``, table scaled by ``ratio`` when decimating;
* ``librosa_resample`` -- librosa/core/audio.py: identity when the rates agree, ``fix_length`` to ``ceil(n * ratio)``.

``tests/test_oracle_dsp.py::test_resample_oracle_pinned_to_torchaudio_kaiser_best`` pins this restatement against
``torchaudio.functional.resample`` configured as torchaudio documents for 'kaiser_best' (a separate implementation of the
same filter): 1e-6 agreement where resampy's integer table stride is exact, and at 48 k -> 16 k -- where resampy truncates the
stride 170.67 -> 170 -- agreement of the un-truncated variant, i.e. the truncation is the only (faithfully restated)
difference.  The same file also checks properties the algorithm must have (unit DC gain, pass-band sinusoids land on
the analytic 16 kHz sinusoid, stop-band tones are removed, agreement with ``scipy.signal.resample_poly`` at the
level two different anti-aliasing filters can agree) -- plausibility pins, not parity pins.
"""
from __future__ import annotations

import numpy as np

KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


def sinc_window(num_zeros: int, precision: int, rolloff: float, beta: float):
    """resampy.filters.sinc_window with ``window = kaiser(beta)``: (half window [n+1], table samples per zero
    crossing, rolloff)."""
    from scipy.signal.windows import kaiser
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits, rolloff


_FILTER_CACHE = {}


def kaiser_best():
    if "kb" not in _FILTER_CACHE:
        _FILTER_CACHE["kb"] = sinc_window(**KAISER_BEST)
    win, num_table, rolloff = _FILTER_CACHE["kb"]
    return win.copy(), num_table, rolloff


def filter_tables(sample_ratio: float):
    """(interp_win, interp_delta, num_table) as resampy.core.resample prepares them."""
    interp_win, num_table, _ = kaiser_best()
    if sample_ratio < 1:
        interp_win *= sample_ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    return interp_win, interp_delta, num_table


def resample_f(x: np.ndarray, n_out: int, sample_ratio: float, interp_win, interp_delta, num_table: int) -> np.ndarray:
    """resampy.interpn.resample_f for one channel, vectorised over the output index (same arithmetic per sample;
    the time register is accumulated by repeated float64 addition exactly as the scalar loop does)."""
    x = np.asarray(x, dtype=np.float64)
    scale = min(1.0, sample_ratio)
    time_increment = 1.0 / sample_ratio
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = x.shape[0]
    # time_register after t additions (np.cumsum adds sequentially in float64, like the loop)
    treg = np.concatenate([[0.0], np.cumsum(np.full(max(n_out - 1, 0), time_increment, dtype=np.float64))])[:n_out]
    n = treg.astype(np.int64)
    y = np.zeros(n_out, dtype=np.float64)
    # left wing
    frac = scale * (treg - n)
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    i_max = np.minimum(n + 1, (nwin - offset) // index_step)
    for i in range(int(i_max.max(initial=0))):
        m = i < i_max
        idx = offset[m] + i * index_step
        y[m] += (interp_win[idx] + eta[m] * interp_delta[idx]) * x[n[m] - i]
    # right wing
    frac = scale - frac
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    k_max = np.minimum(n_orig - n - 1, (nwin - offset) // index_step)
    for k in range(int(k_max.max(initial=0))):
        m = k < k_max
        idx = offset[m] + k * index_step
        y[m] += (interp_win[idx] + eta[m] * interp_delta[idx]) * x[n[m] + k + 1]
    return y


def resampy_resample(x: np.ndarray, sr_orig: int, sr_new: int) -> np.ndarray:
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_best') for a 1-D signal."""
    sample_ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[-1] * sample_ratio)
    if n_out < 1:
        raise ValueError("input too short to resample")
    interp_win, interp_delta, num_table = filter_tables(sample_ratio)
    return resample_f(x, n_out, sample_ratio, interp_win, interp_delta, num_table)


def librosa_resample(y: np.ndarray, orig_sr: int, target_sr: int, fix: bool = True, scale: bool = False) -> np.ndarray:
    """librosa.resample(y, orig_sr, target_sr, fix=True, scale=False) of librosa <= 0.7 (res_type='kaiser_best')."""
    y = np.asarray(y, dtype=np.float64)
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    y_hat = resampy_resample(y, orig_sr, target_sr)
    if fix:                                                   # librosa.util.fix_length: zero-pad or trim the tail
        if y_hat.shape[-1] < n_samples:
            y_hat = np.concatenate([y_hat, np.zeros(n_samples - y_hat.shape[-1])])
        else:
            y_hat = y_hat[:n_samples]
    if scale:
        y_hat = y_hat / np.sqrt(ratio)
    return np.ascontiguousarray(y_hat, dtype=y.dtype)
