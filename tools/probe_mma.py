"""tcgen05.mma microbenchmark: cycles per MMA for issue styles, N, aligned vs shifted descriptors."""
import os as _os; _os.environ["SG2_PROBES"] = "1"   # diagnostics build: python -m sg2b200.build --probes
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("style N rowB shift pitch nacc kpi mn blocks | issue_cyc/mma total_cyc/mma  (floor = N/2)")
iters = 256
def run(style, N, rb, shift, pitch, nacc, kpi, mn, nb):
    _lib.call("sg2_probe_mma", out.data_ptr(), N, rb, shift, pitch, nacc, iters, kpi, mn, style, nb, st)
    torch.cuda.synchronize()
    a, b = out.tolist()
    n = iters * kpi
    print(f"{style} {N:4d} {rb:4d} {shift:3d} {pitch:3d} {nacc:2d} {kpi:2d} {mn:1d} {nb:4d} | {a/n:8.1f} {b/n:8.1f}", flush=True)
for style in (0, 1, 2, 3):
    for N in (16, 32, 64, 128, 256):
        run(style, N, 128, 0, 8, 1, 4, 0, 148)
        run(style, N, 128, 11, 10, 1, 4, 0, 148)
    run(style, 32, 64, 11, 10, 1, 4, 0, 148)
    run(style, 32, 32, 11, 10, 1, 4, 0, 148)
    run(style, 64, 128, 11, 10, 1, 4, 1, 148)
