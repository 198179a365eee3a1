"""Is a model's decode loop bound by the host (Python + ctypes + launch) or by the GPU?  For each model: wall time until
the enhance call RETURNS (everything enqueued) vs wall time until the device is idle.  enqueue ~= total means the GPU
waits for the host.  Development tool."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_models as bm          # noqa: E402
import se_b200                     # noqa: E402
from oracle import synth           # noqa: E402


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or list(bm.CASES)
    dev = torch.device("cuda")
    for name in which:
        ctor, tmpl, genh, oenh, bsz, secs, hop, kw = bm.CASES[name]
        if name == "ctsnet":
            sd = tuple(synth.synthetic_state_dict(t, seed=i, gain=1.0) for i, t in enumerate(tmpl()))
            model = ctor()
            for m, s in zip(model, sd):
                m.load_state_dict(s)
                m.eval().cuda()
        else:
            model = ctor()
            model.load_state_dict(synth.synthetic_state_dict(tmpl(), seed=0, gain=1.0))
            model.eval().cuda()
        wav = torch.from_numpy(synth.noisy_batch(bsz, 16000 * secs)).to(dev)
        for _ in range(2):
            genh(model, wav, **kw)
        torch.cuda.synchronize()
        n0 = se_b200.ops.launch_count()
        t0 = time.perf_counter()
        genh(model, wav, **kw)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rec = {"model": name, "enqueue_ms": 1e3 * (t1 - t0), "total_ms": 1e3 * (t2 - t0),
               "launches": se_b200.ops.launch_count() - n0,
               "host_us_per_launch": 1e6 * (t1 - t0) / max(1, se_b200.ops.launch_count() - n0)}
        # the same batch through decode.GraphedEnhance: one graph launch
        y_eager = genh(model, wav, **kw)
        dec = se_b200.decode.GraphedEnhance(model, genh, **kw)
        dec(wav)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        yg = dec(wav)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rec.update(graph_enqueue_ms=1e3 * (t1 - t0), graph_total_ms=1e3 * (t2 - t0),
                   graph_enqueue_share=(t1 - t0) / (t2 - t0), graph_equals_eager=bool(torch.equal(yg, y_eager)))
        print(json.dumps(rec), flush=True)
        del dec


if __name__ == "__main__":
    main()
