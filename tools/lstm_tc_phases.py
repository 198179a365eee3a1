"""Phase breakdown of the tcgen05 LSTM recurrence (csrc/lstm_tc.cu): runs one H=1024, B=64 sequence with
the clock64() stamps switched on (se_debug_lstm_tc_profile) and prints where a step's time goes.
Usage (GPU box): python tools/lstm_tc_phases.py [T] > gpurun_out/lstm_tc_phases.txt"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
from se_b200 import _lib, ops  # noqa: E402

EV = ["poll_start", "barrier_seen", "tma_issued", "first_stage", "last_stage", "acc_done", "dsmem_sent",
      "cluster_passed", "cell_done", "arrived", "proxy_fenced", "cta_synced"]
# engine 4 (csrc/lstm_f16.cu)
EV4 = ["load_start", "first_item_valid", "tiles_handed_over", "mma_saw_chunk0", "mma_committed", "acc_done", "dsmem_sent",
       "cluster_passed", "published", "step_done", "-", "-"]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 401
    engine = int(sys.argv[2]) if len(sys.argv) > 2 else 4          # 2 = csrc/lstm_tc.cu, 4 = csrc/lstm_f16.cu
    B, H, NS = 64, 1024, 8
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    xp = torch.randn(B, T, 4 * H, generator=g).to(dev)
    whh = (torch.randn(H // 8, H, 32, generator=g) / 32).to(dev)
    hs = torch.empty(B, T, H, device=dev)
    ops.set_lstm_engine(engine)
    for _ in range(2):
        ops.lstm_seq(xp, whh, H, out=hs)
    torch.cuda.synchronize()
    buf = torch.zeros(128 * NS * len(EV), dtype=torch.int64, device=dev)
    t0 = T // 2
    lib = _lib.load()
    assert lib.se_debug_lstm_tc_profile(ctypes.c_void_p(buf.data_ptr()), t0, NS) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.lstm_seq(xp, whh, H, out=hs)
    e1.record()
    torch.cuda.synchronize()
    lib.se_debug_lstm_tc_profile(None, 1, 1)
    ms = e0.elapsed_time(e1)
    st = buf.cpu().numpy().reshape(128, NS, len(EV)).astype(np.float64)
    # SM clocks are not synchronised across SMs: only per-CTA differences are meaningful
    step = np.diff(st[:, :, 9], axis=1).mean()            # arrival -> arrival, cycles
    out = {"engine": engine, "T": T, "ms": ms, "us_per_step": 1e3 * ms / T, "cycles_per_step": step,
           "ghz_implied": step / (1e3 * ms / T) / 1e3}
    seg = {}
    if engine >= 4:
        names, pairs = EV4, [(0, 1), (1, 2), (1, 3), (3, 4), (2, 5), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9)]
        step = np.diff(st[:, :, 8], axis=1).mean()        # publish -> publish
        out["cycles_per_step"] = step
        out["ghz_implied"] = step / (1e3 * ms / T) / 1e3
    else:
        names, pairs = EV, [(0, 1), (1, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (8, 10), (10, 11), (11, 9), (1, 2)]
    for a, b in pairs:
        d = st[:, 1:, b] - st[:, 1:, a]
        seg[f"{names[a]}->{names[b]}"] = {"mean": float(d.mean()), "p10": float(np.percentile(d, 10)),
                                          "p90": float(np.percentile(d, 90)), "max": float(d.max())}
    # previous step's publish / arrival -> this step's first valid item / barrier seen (includes waiting for the
    # slowest producer)
    a, b = (8, 1) if engine >= 4 else (9, 1)
    d = st[:, 1:, b] - st[:, :-1, a]
    seg[f"{names[a]}(t-1)->{names[b]}(t)"] = {"mean": float(d.mean()), "p10": float(np.percentile(d, 10)),
                                              "p90": float(np.percentile(d, 90)), "max": float(d.max())}
    out["segments_cycles"] = seg
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
