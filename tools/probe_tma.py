"""TMA streaming-bandwidth microbenchmark: GB/s vs row bytes (channels), box shape, ring depth, CTAs per SM."""
import os as _os; _os.environ["SG2_PROBES"] = "1"   # diagnostics build: python -m sg2b200.build --probes
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("C rowB  box(w x h) depth ctas swz | MB  us  GB/s")
def run(C, H, bw, bh, depth, nb, swz, B=24):
    x = torch.empty(B, H, H, C, device="cuda", dtype=torch.bfloat16).normal_()
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    _lib.call("sg2_probe_tma", x.data_ptr(), out.data_ptr(), B, H, H, C, bw, bh, depth, nb, swz, st)
    e1.record()
    torch.cuda.synchronize()
    us = 1000 * e0.elapsed_time(e1)
    mb = x.numel() * 2 / 1e6
    print(f"{C:3d} {2*C:4d}  {bw:3d}x{bh:<3d} {depth:3d} {nb:4d} {swz:4d} | {mb:6.1f} {us:8.1f} {mb/us*1e3:8.1f}", flush=True)
for C, H in ((16, 256), (32, 128), (64, 128), (64, 64)):
    swz = 2 * C
    for (bw, bh) in ((8, 16), (16, 8), (32, 4), (16, 16)):
        for depth in (2, 4, 8):
            for nb in (148, 296, 592):
                if depth * bw * bh * 2 * C * (nb // 148) > 200 * 1024:
                    continue
                run(C, H, bw, bh, depth, nb, swz)
