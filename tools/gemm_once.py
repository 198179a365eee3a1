"""One warm + a few timed launches of the CRN projection GEMM (25664 x 1024 x 4096, 3xTF32) on the engine selected by
SE_GEMM_ENGINE: the target of `ncu --set full -k regex:gemm_tf32x3`.  Development tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
M, K, N = 25664, 1024, 4096
x_hi, x_lo = ops.split_tf32(torch.randn(M, K, generator=g).to(dev))
w_hi, w_lo = ops.split_tf32((torch.randn(N, K, generator=g) / 32).to(dev))
bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.gemm_tf32x3(x_hi, x_lo, w_hi, w_lo, bias, N, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0]))
