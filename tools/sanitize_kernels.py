"""Small launches of the kernels with hand-rolled synchronisation, for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py
    compute-sanitizer --tool memcheck  python tools/sanitize_kernels.py
istft_kernel (shared-memory gather overlap-add), stft_kernel (bulk-copy staged tile), lstm_seq_f16_kernel (DSMEM K-split
reduction, tagged-state exchange), lstm_seq_tc_kernel (engine 2), lstm_seq_small_kernel (H = 128), lstm_seq_kernel (FMA
engine, device-wide barrier).  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200                                    # noqa: E402
from se_b200 import ops                           # noqa: E402
from se_b200._lib import ISTFT_SPEC               # noqa: E402
from oracle import synth                          # noqa: E402


def main():
    dev = torch.device("cuda")
    which = sys.argv[1:] or ["dsp", "lstm_f16", "lstm_tc", "lstm_small", "lstm_fma"]
    if "dsp" in which:
        for n_fft, win, hop in ((320, 320, 160), (512, 400, 160)):
            wav = torch.from_numpy(synth.noisy_batch(2, 6400 + 37)).to(dev)
            lens = torch.tensor([6437, 5000], dtype=torch.int32, device=dev)
            c, ic = ops.rms_scale(wav, lengths=lens)
            t, f = 1 + wav.shape[1] // hop, n_fft // 2 + 1
            spec = torch.empty(2, t, f, 2, device=dev)
            ops.stft(wav, c, n_fft, win, hop, re=spec[..., 0], im=spec[..., 1], lengths=lens)
            out = torch.empty_like(wav)
            ops.istft(ISTFT_SPEC, spec[..., 0], spec[..., 1], None, None, n_fft, win, hop, out, wav.shape[1], out_scale=ic,
                      lengths=lens)
            torch.cuda.synchronize()
            print("dsp", n_fft, hop, float((out[0] - wav[0]).abs().max()))
    g = torch.Generator().manual_seed(0)
    for name, eng, h, b, t in (("lstm_f16", 4, 1024, 64, 4), ("lstm_tc", 2, 1024, 64, 3), ("lstm_small", 3, 128, 5, 6),
                               ("lstm_fma", 0, 512, 3, 3)):
        if name not in which:
            continue
        ops.set_lstm_engine(eng)
        xp = torch.randn(b, t, 4 * h, generator=g).to(dev)
        whh = (torch.randn(h // 8, h, 32, generator=g) / np.sqrt(h)).to(dev)
        y = ops.lstm_seq(xp, whh, h)
        torch.cuda.synchronize()
        print(name, float(y.abs().max()))
    ops.set_lstm_engine(4)


if __name__ == "__main__":
    main()
