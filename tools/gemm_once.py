"""One warm + a few timed launches of the CRN projection GEMM (25664 x 1024 x 4096) on the engine selected by
SE_GEMM_ENGINE: the target of `ncu --set full -k regex:gemm_tf32x3`.  `f16` as second argument runs the fp16-pair engine
(se_gemm_f16x3), anything else the 3xTF32 one.  Development tool."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se_b200 import ops, packing  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
M, K, N = 25664, 1024, 4096
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
f16 = len(sys.argv) > 2 and sys.argv[2] == "f16"
x = torch.randn(M, K, generator=g).to(dev)
w = (torch.randn(N, K, generator=g) / 32).to(dev)
bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev)
if f16:
    a = ops.split_f16(x)
    w_hi, w_lo, ws = packing.split_f16(w)
    run = lambda: ops.gemm_f16x3(a, (w_hi, w_lo), ws, bias, N, out=out)   # noqa: E731
else:
    x_hi, x_lo = ops.split_tf32(x)
    w_hi, w_lo = ops.split_tf32(w)
    run = lambda: ops.gemm_tf32x3(x_hi, x_lo, w_hi, w_lo, bias, N, out=out)   # noqa: E731
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{'f16x3' if f16 else 'tf32x3'} {M}x{K}x{N}: {ms:.3f} ms per launch = {2 * M * K * N / ms / 1e9:.1f} TFLOP/s fp32-equivalent "
      f"({3 * 2 * M * K * N / ms / 1e9:.1f} tensor), out[0,0] = {float(out[0, 0]):.6f}")
