"""Frames/s of the other BASELINE.json configs on one GPU (per-GPU shard sizes of configs 3 and 4,
plus the LSTM model at the CRN batch).  Development/report tool; bench.py stays on configs[1].
    python tools/bench_models.py [--cpu]   -> JSON lines
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200                                       # noqa: E402
from oracle import decode as odecode                 # noqa: E402
from oracle import synth, templates                  # noqa: E402

FSN_ARGS = dict(num_freqs=257, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
                fb_output_activate_function="ReLU", sb_output_activate_function=None, fb_model_hidden_size=512,
                sb_model_hidden_size=384)

CASES = {
    # name: (ctor, template, gpu enhance, oracle enhance, batch per GPU, seconds, hop, kwargs)
    "lstm": (lambda: se_b200.lstm_net(), templates.lstm_template, se_b200.decode.enhance_lstm, odecode.enhance_lstm,
             64, 4, 160, dict(p=1.0)),
    "dccrn": (lambda: se_b200.DCCRN(rnn_units=256, masking_mode='E', use_clstm=True,
                                    kernel_num=[32, 64, 128, 256, 256, 256]), templates.dccrn_template,
              se_b200.decode.enhance_dccrn, odecode.enhance_dccrn, 32, 4, 128, dict(p=0.5)),
    "fullsubnet": (lambda: se_b200.fullsubnet.Model(**FSN_ARGS), templates.fullsubnet_template,
                   se_b200.decode.enhance_fullsubnet, odecode.enhance_fullsubnet, 32, 10, 256, dict(p=0.5)),
    "gcrn": (lambda: se_b200.gcrn.Net(), templates.gcrn_template, se_b200.decode.enhance_gcrn, odecode.enhance_gcrn,
             64, 4, 160, dict(p=0.5)),
    "dpcrn": (lambda: se_b200.dpcrn(), templates.dpcrn_template, se_b200.decode.enhance_dpcrn, odecode.enhance_dpcrn,
              64, 4, 160, dict(p=1.0)),
    "ctsnet": (lambda: (se_b200.ctsnet.Step1_net(), se_b200.ctsnet.Step2_net(X=6, R=3)),
               lambda: (templates.ctsnet_step1_template(), templates.ctsnet_step2_template()),
               se_b200.decode.enhance_ctsnet, odecode.enhance_ctsnet, 64, 4, 160, dict(p=1.0)),
    "taylor": (lambda: se_b200.TaylorSENet(), templates.taylorsenet_template, se_b200.decode.enhance_taylorsenet,
               odecode.enhance_taylorsenet, 64, 4, 160, dict(p=1.0)),
    "g2net": (lambda: se_b200.g2net.gaf_base(3, 64, 2, 4, 4, [1, 2, 5, 9], 256 + 161 * 2, 256, 256, (2, 3), (1, 3), 64, 'cat',
                                             3, is_aux=False, encoder_type='U2Net', tcm_type='full-band'),
              templates.g2net_template, se_b200.decode.enhance_g2net, odecode.enhance_g2net, 64, 4, 160, dict(p=0.5)),
    "uformer": (lambda: se_b200.Uformer(), templates.uformer_template, se_b200.decode.enhance_uformer, None,
                64, 4, 160, dict()),
}


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    do_cpu = "--cpu" in sys.argv
    dev = torch.device("cuda")
    for name in which:
        ctor, tmpl, genh, oenh, bsz, secs, hop, kw = CASES[name]
        if name == "ctsnet":                      # two stages: tuples of templates / state-dicts / modules
            sd = tuple(synth.synthetic_state_dict(t, seed=i, gain=1.0) for i, t in enumerate(tmpl()))
            model = ctor()
            for m, s in zip(model, sd):
                m.load_state_dict(s)
                m.eval().cuda()
        else:
            sd = synth.synthetic_state_dict(tmpl(), seed=0, gain=1.0 if name in ("uformer", "dpcrn", "taylor", "g2net") else 2.0)
            model = ctor()
            model.load_state_dict(sd)
            model.eval().cuda()
        n = 16000 * secs
        frames = 1 + n // hop
        base = synth.noisy_batch(min(bsz, 8), n)
        wav = torch.from_numpy(np.concatenate([base] * (bsz // base.shape[0]), axis=0)).to(dev)
        genh(model, wav, **kw)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            se_b200.ops.start_recording()
            e0.record()
            genh(model, wav, **kw)
            e1.record()
            torch.cuda.synchronize()
            per_op = se_b200.ops.stop_recording()
            times.append(e0.elapsed_time(e1))
        ms = min(times)
        share = {k: round(v[1] / sum(x[1] for x in per_op.values()), 3)
                 for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][1])[:(64 if "--all-ops" in sys.argv else 8)]}
        if "--all-ops" in sys.argv:      # launches per op as well
            share = {k: [v, per_op[k][0]] for k, v in share.items()}
        rec = {"model": name, "batch": bsz, "clip_s": secs, "frames_per_clip": frames, "ms_per_batch": ms,
               "frames_per_s": bsz * frames / (ms * 1e-3), "rtf": ms * 1e-3 / (bsz * secs), "op_time_share": share,
               "weights": "seeded synthetic"}
        if do_cpu and oenh is not None:
            torch.set_num_threads(16)
            x = base[0].astype(np.float64)
            oenh(sd, x, **kw)
            t0 = time.perf_counter()
            oenh(sd, x, **kw)
            dt = time.perf_counter() - t0
            rec["cpu_frames_per_s_16thr"] = frames / dt
        print(json.dumps(rec), flush=True)
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
