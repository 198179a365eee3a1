"""Host enqueue time vs GPU time of one CRN decode step (64 x 4 s): is the eager loop host-bound?  Development tool."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200                                       # noqa: E402
from oracle import synth, templates                  # noqa: E402

sd = synth.synthetic_state_dict(templates.crn_template(), seed=0)
m = se_b200.crn_net()
m.load_state_dict(sd)
m.eval().cuda()
wav = torch.from_numpy(synth.noisy_batch(64, 64000)).cuda()
for _ in range(3):
    se_b200.decode.enhance_crn(m, wav)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    se_b200.decode.enhance_crn(m, wav)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"eager: enqueue {1e3 * (t1 - t0) / n:.2f} ms per step, until idle {1e3 * (t2 - t0) / n:.2f} ms per step")
g = se_b200.decode.GraphedEnhance(m)
for _ in range(3):
    g(wav)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    g(wav)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"graph: enqueue {1e3 * (t1 - t0) / n:.2f} ms per step, until idle {1e3 * (t2 - t0) / n:.2f} ms per step")
