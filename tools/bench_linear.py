"""Graph-replay timing of the INIT_STAGE_G fc (228 -> 32768, M = 24) linear kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sg2b200 import ops
from layer_bench import timed
dev = torch.device("cuda:0")
M, K1, K2, N = 24, 128, 100, 32768
c, z = torch.randn(M, K1, device=dev), torch.randn(M, K2, device=dev)
w = torch.randn(N, K1 + K2, device=dev) * 0.01
dy = torch.randn(M, N, device=dev).bfloat16()
dw = torch.empty(N, K1 + K2, device=dev)
flush = torch.empty(64 << 20, device=dev)
for name, fn in (("linear_fwd", lambda: ops.linear_fwd(c, z, w, None, False)),
                 ("linear_bwd_w", lambda: ops.linear_bwd_w(dy, c, z, dw, None, False)),
                 ("linear_bwd_x", lambda: ops.linear_bwd_x(dy, w, K1))):
    print(f"{name:14s} {timed(fn):8.1f} us (L2-warm)   {timed(lambda: (flush.zero_(), fn())):8.1f} us incl. a 256 MB flush fill")
print(f"flush alone    {timed(lambda: flush.zero_()):8.1f} us")
