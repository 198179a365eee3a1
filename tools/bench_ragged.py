"""Throughput of length-bucketed decode on a VoiceBank-like length distribution.
824 clips (the size of the VoiceBank+DEMAND test set), lengths drawn log-normally around 2.5 s and clipped to
[1.1 s, 9.8 s] (that set's range), 16 kHz, in-memory (no disk): decode.plan_batches + per-clip lengths against (a) the
reference's own schedule, one file at a time (B = 1), and (b) the equal-length best case.  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200                                       # noqa: E402
from oracle import synth, templates                  # noqa: E402


def main():
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    nclips = int(sys.argv[1]) if len(sys.argv) > 1 else 824
    lens = np.clip(np.exp(rng.normal(np.log(2.5), 0.45, nclips)), 1.1, 9.8)
    lens = (lens * 16000).astype(int)
    ckpt = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "checkpoints", "_ref",
                        "CRN__wsj0_si84_300h_crn_noncprs_model.pth")
    sd = torch.load(ckpt, map_location="cpu") if os.path.exists(ckpt) else synth.synthetic_state_dict(templates.crn_template(), seed=0)
    model = se_b200.crn_net()
    model.load_state_dict(sd)
    model.eval().cuda()
    base = synth.noisy_clip(3, int(lens.max()))
    infos = [(f"{i}", 16000, int(n)) for i, n in enumerate(lens)]
    plan = se_b200.decode.plan_batches(infos, 64, True, 0.25)
    frames = int(sum(1 + n // 160 for n in lens))

    def run_bucketed():
        for _, names, ls in plan:
            nmax = max(ls)
            wav = torch.zeros(len(ls), nmax, device=dev)
            for i, n in enumerate(ls):
                wav[i, :n] = torch.from_numpy(np.roll(base, 17 * i)[:n]).to(dev)
            se_b200.decode.enhance_crn(model, wav, lengths=ls)

    def run_b1(k):
        for n in lens[:k]:
            se_b200.decode.enhance_crn(model, torch.from_numpy(base[:n].copy()).to(dev)[None])

    run_bucketed()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_bucketed()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    k = 64
    run_b1(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_b1(k)
    torch.cuda.synchronize()
    dt1 = time.perf_counter() - t0
    padded = sum(len(ls) * max(ls) - sum(ls) for _, _, ls in plan) / sum(len(ls) * max(ls) for _, _, ls in plan)
    print(json.dumps({
        "workload": f"CRN decode of {nclips} clips, VoiceBank-like lengths {lens.min() / 16000:.1f}..{lens.max() / 16000:.1f} s "
                    f"(mean {lens.mean() / 16000:.2f} s), in memory",
        "batches": len(plan), "padded_fraction": padded, "frames": frames,
        "bucketed_frames_per_s": frames / dt, "bucketed_s": dt,
        "one_file_at_a_time_frames_per_s": float(sum(1 + n // 160 for n in lens[:k]) / dt1),
        "speedup_vs_b1": (frames / dt) / (sum(1 + n // 160 for n in lens[:k]) / dt1),
        "note": "wall clock incl. host-side batch assembly; B = 1 is the reference's schedule (crn_decode_vb.py:31)"}))


if __name__ == "__main__":
    main()
