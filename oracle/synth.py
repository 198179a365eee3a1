"""Deterministic synthetic inputs and weights shared by the tests, the goldens and bench.py.

TEST / BENCH INFRASTRUCTURE.  No reference code corresponds to this file: the
reference decodes real corpora with shipped checkpoints.  There is no dataset here,
so clips follow SURVEY.md section 8(d) ("noisy speech" surrogate) and, where the shipped
checkpoints cannot travel, weights are drawn per state-dict key from a seed derived
from the key name -- independent of module construction order, so the reference
module, the oracle and the CUDA path can all be given bit-identical parameters.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch
from scipy.signal import lfilter

FS = 16000
SNR_GRID = (-5.0, 0.0, 5.0, 10.0)   # the reference's own test grid, lstm_decode.py:75-...


def noisy_clip(index: int, num_samples: int, base_seed: int = 1234) -> np.ndarray:
    """Speech-like harmonic source with syllabic AM and gaps + low-passed Gaussian noise.
    float32, peak 0.5.  Clip ``index`` uses seed ``base_seed + index``."""
    rng = np.random.default_rng(base_seed + index)
    n = num_samples
    t = np.arange(n) / FS
    # slow random-walk F0 contour in 100..250 Hz
    steps = rng.standard_normal(n // 160 + 2) * 4.0
    f0_frames = np.clip(150.0 + np.cumsum(steps), 100.0, 250.0)
    f0 = np.interp(np.arange(n), np.arange(len(f0_frames)) * 160, f0_frames)
    phase = 2.0 * np.pi * np.cumsum(f0) / FS
    speech = np.zeros(n)
    for h in range(1, 11):
        speech += np.sin(h * phase + rng.uniform(0, 2 * np.pi)) / h
    am_rate = rng.uniform(3.0, 5.0)
    am = 0.5 - 0.5 * np.cos(2.0 * np.pi * am_rate * t + rng.uniform(0, 2 * np.pi))
    am = am ** 2
    # ~30 % silent gaps, in 200 ms blocks
    blocks = rng.uniform(size=n // 3200 + 1) > 0.3
    gate = np.repeat(blocks.astype(np.float64), 3200)[:n]
    speech = speech * am * gate
    noise = rng.standard_normal(n)
    a = 0.7
    lp = lfilter([1.0 - a], [1.0, -a], noise)   # one-pole low-pass
    snr = SNR_GRID[index % len(SNR_GRID)]
    ps = np.mean(speech ** 2) + 1e-12
    pn = np.mean(lp ** 2) + 1e-12
    mix = speech + lp * np.sqrt(ps / (pn * 10.0 ** (snr / 10.0)))
    mix = 0.5 * mix / (np.max(np.abs(mix)) + 1e-12)
    return mix.astype(np.float32)


def noisy_batch(batch: int, num_samples: int, first_index: int = 0) -> np.ndarray:
    return np.stack([noisy_clip(first_index + i, num_samples) for i in range(batch)])


def _key_seed(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def synthetic_state_dict(template: dict, seed: int = 0, gain: float = 2.0) -> dict:
    """Fill a state-dict with the shapes/keys of ``template`` (name -> tensor or shape tuple).

    conv / linear / lstm weights: U(-k, k) with k = 1/sqrt(fan_in) (torch's default scale);
    BatchNorm: weight U(0.5,1.5), bias U(-0.2,0.2), running_mean N(0,0.2), running_var U(0.5,1.5).
    ``gain`` multiplies every >=2-D weight so that signals neither die out nor saturate (gain 2
    gives CRN outputs with the dynamic range of the shipped checkpoints).
    """
    out = {}
    for name, ref in template.items():
        shape = tuple(ref.shape) if hasattr(ref, "shape") else tuple(ref)
        g = torch.Generator().manual_seed(_key_seed(name, seed))
        if name.endswith("num_batches_tracked"):
            out[name] = torch.tensor(1000, dtype=torch.int64)
        elif len(shape) == 1 and shape[0] == 1 and name.endswith(".2.weight") and not _looks_like_norm(name, template):
            out[name] = torch.full(shape, 0.25) + 0.1 * (torch.rand(shape, generator=g) - 0.5)   # nn.PReLU slope
        elif name.endswith("running_var"):
            out[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            out[name] = torch.randn(shape, generator=g) * 0.2
        elif len(shape) == 1 and (".weight" in name) and ("lstm" not in name) and _looks_like_norm(name, template):
            out[name] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1 and name.endswith(".bias") and _looks_like_norm(name, template):
            out[name] = (torch.rand(shape, generator=g) - 0.5) * 0.4
        else:
            if len(shape) >= 2:
                fan_in = int(np.prod(shape[1:]))
                if "lstm" in name or "rnn" in name:
                    fan_in = shape[0] // 4  # torch LSTM init uses 1/sqrt(hidden)
            else:
                fan_in = max(shape[0], 1) if shape else 1
                if "lstm" in name:
                    fan_in = shape[0] // 4
            k = 1.0 / np.sqrt(max(fan_in, 1))
            out[name] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * k * (gain if len(shape) >= 2 else 1.0)
    return out


def _looks_like_norm(name: str, template: dict) -> bool:
    stem = name.rsplit(".", 1)[0]
    return (stem + ".running_mean") in template
