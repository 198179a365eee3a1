"""DSP-only run of SURVEY.md section 8(d): rms-scale -> STFT -> (identity) -> iSTFT on 4 s clips for the four
geometries of the reference (512/512/256 is the literal "512-FFT / 256-hop" of BASELINE.json's metric string), per
kernel: time (CUDA events, best of N, inputs rotated over more than L2), frames/s, algorithmic bytes / time and the
fraction of the measured HBM peak.  Development / report tool -> JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200                                    # noqa: E402
from se_b200 import ops                           # noqa: E402
from se_b200._lib import ISTFT_SPEC               # noqa: E402
from oracle import synth                          # noqa: E402

GEOMS = {"320/320/160": (320, 320, 160), "512/512/128": (512, 512, 128), "512/512/256": (512, 512, 256),
         "512/400/160": (512, 400, 160)}


def best_of(fn_list, iters=4):
    """fn_list: one closure per rotated input set; returns the best per-call time in ms over `iters` sweeps."""
    for fn in fn_list:
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for fn in fn_list:
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(fn_list))
    return best


def main():
    dev = torch.device("cuda")
    peak = 6532.2
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    n = 64000
    for bsz in (64, 512):
        base = torch.from_numpy(synth.noisy_batch(min(bsz, 16), n))
        wav0 = torch.cat([base] * (bsz // base.shape[0]), 0).to(dev)
        nsets = max(2, int(200e6 // (bsz * n * 4)) + 1)                 # rotate inputs over > 126 MB of L2
        wavs = [torch.roll(wav0, 997 * i, dims=1).contiguous() for i in range(nsets)]
        for name, (n_fft, win, hop) in GEOMS.items():
            t, f = 1 + n // hop, n_fft // 2 + 1
            c, inv_c = ops.rms_scale(wavs[0])
            specs = [torch.empty(bsz, t, f, 2, device=dev) for _ in range(nsets)]
            outs = [torch.empty(bsz, n, device=dev) for _ in range(nsets)]
            ms_s = best_of([(lambda w=w, s=s: ops.stft(w, c, n_fft, win, hop, re=s[..., 0], im=s[..., 1]))
                            for w, s in zip(wavs, specs)])
            ms_i = best_of([(lambda s=s, o=o: ops.istft(ISTFT_SPEC, s[..., 0], s[..., 1], None, None, n_fft, win, hop, o, n,
                                                        out_scale=inv_c)) for s, o in zip(specs, outs)])
            err = (outs[0] - wavs[0]).abs().max().item()
            frames = bsz * t
            b_stft = frames * (4 * hop + 8 * f)          # audio in + complex spectrum out
            b_istft = frames * (8 * f + 4 * hop)         # complex spectrum in + audio out
            print(json.dumps({
                "geometry": name, "batch": bsz, "frames": frames, "roundtrip_max_abs_err": err,
                "stft_us": 1e3 * ms_s, "stft_frames_per_s": frames / (ms_s * 1e-3), "stft_GBps": b_stft / ms_s / 1e6,
                "stft_frac_of_hbm_peak": b_stft / ms_s / 1e6 / peak,
                "istft_us": 1e3 * ms_i, "istft_frames_per_s": frames / (ms_i * 1e-3), "istft_GBps": b_istft / ms_i / 1e6,
                "istft_frac_of_hbm_peak": b_istft / ms_i / 1e6 / peak,
                "dsp_only_frames_per_s": frames / ((ms_s + ms_i) * 1e-3), "hbm_peak_GBps": peak}), flush=True)
            del specs, outs
        del wavs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
