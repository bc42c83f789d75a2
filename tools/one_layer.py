"""Run one conv op of one layer a few times (for ncu): python tools/one_layer.py G.res3b fprop [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sg2b200 import ops
import layer_bench as lb
name, op = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
for n, kind, H, Cin, Cout, act, cnt in lb.LAYERS:
    if n != name:
        continue
    B = lb.B
    if kind == ops.GEMM:
        Bx, Hx, Wx = 1, 1, B * H * H
    else:
        Bx, Hx, Wx = B, H, H
    x = torch.randn(Bx, Hx, Wx, Cin, device=dev).bfloat16()
    s1, s2 = ops.pack_shapes(kind, Cout, Cin)
    wpk = (torch.randn(s1, device=dev) * 0.02).bfloat16()
    wpkT = (torch.randn(s2, device=dev) * 0.02).bfloat16()
    Ho, Wo = ops._out_hw(kind, Hx, Wx)
    dy = torch.randn(Bx, Ho, Wo, Cout, device=dev).bfloat16()
    dwpk = torch.zeros(Cout, ops.JOBS[kind], Cin, device=dev, dtype=torch.float32)
    st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64) if act is not None else None
    for _ in range(reps):
        if op == "fprop":
            ops.conv_fprop(kind, x, wpk, Cout, stats=st)
        elif op == "dgrad":
            ops.conv_dgrad(kind, dy, wpkT, Bx, Hx, Wx, Cin)
        else:
            ops.conv_wgrad(kind, x, dy, dwpk)
    torch.cuda.synchronize()
    print("ok", name, op)
