"""Pin the oracle's numpy STFT/iSTFT restatement (librosa algorithm) against torch.stft/istft --
the operators the torch-dialect decode scripts call directly (DCCRN/dccrn_decode.py:41,56)."""
import numpy as np
import pytest
import torch

from oracle import dsp, synth


@pytest.mark.parametrize("geom", list(dsp.GEOMETRIES))
@pytest.mark.parametrize("n", [16000, 16000 + 37, 4000])
def test_stft_matches_torch(geom, n):
    n_fft, win, hop = dsp.GEOMETRIES[geom]
    x = synth.noisy_clip(3, n).astype(np.float64)
    s = dsp.stft(x, n_fft, win, hop, out_dtype=np.complex128)
    st = torch.stft(torch.from_numpy(x), n_fft, hop, win, torch.hann_window(win, dtype=torch.float64),
                    return_complex=True).numpy()
    assert s.shape == (n_fft // 2 + 1, dsp.num_frames(n, hop))
    assert np.abs(s - st).max() < 1e-10


@pytest.mark.parametrize("geom", list(dsp.GEOMETRIES))
@pytest.mark.parametrize("length", ["n", None])
def test_istft_matches_torch(geom, length):
    n_fft, win, hop = dsp.GEOMETRIES[geom]
    n = 16000
    x = synth.noisy_clip(4, n).astype(np.float64)
    s = dsp.stft(x, n_fft, win, hop)
    L = n if length == "n" else None
    y = dsp.istft(s, n_fft, win, hop, L)
    yt = torch.istft(torch.from_numpy(s), n_fft, hop, win, torch.hann_window(win), length=L).numpy()
    assert y.shape == yt.shape
    assert np.abs(y - yt).max() < 5e-6
    m = min(len(y), n)
    assert np.abs(y[:m] - x[:m]).max() < 5e-6      # perfect reconstruction property


def test_rms_scale_and_length_rules():
    x = synth.noisy_clip(0, 8000).astype(np.float64)
    xs, c = dsp.rms_scale(x)
    assert abs(np.sqrt(np.mean(xs ** 2)) - 1.0) < 1e-12
    # librosa fix_length pads with zeros when the overlap-add is shorter than `length`
    s = dsp.stft(xs, 320, 320, 160)
    y = dsp.istft(s, 320, 320, 160, length=9000)
    assert len(y) == 9000 and np.all(y[8160:] == 0)
