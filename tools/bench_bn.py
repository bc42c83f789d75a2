"""Time the BN(+GLU/LReLU) forward / backward kernels on the step's big tensors against the HBM roofline.

    python tools/bench_bn.py            # table: shape, act, fwd us / GB/s, bwd us / GB/s  (algorithmic bytes)
Buffers rotate over several copies so every launch reads from HBM, not from the 126 MB L2.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import ops

dev = torch.device("cuda:0")
CASES = [  # (rows per group, C, act, groups)
    (24 * 256 * 256, 32, ops.ACT_GLU, 1), (24 * 128 * 128, 128, ops.ACT_GLU, 1), (24 * 128 * 128, 64, ops.ACT_GLU, 1),
    (24 * 128 * 128, 32, ops.ACT_NONE, 1), (24 * 64 * 64, 256, ops.ACT_GLU, 1), (24 * 64 * 64, 128, ops.ACT_GLU, 1),
    (24 * 64 * 64, 64, ops.ACT_NONE, 1), (24 * 32 * 32, 512, ops.ACT_GLU, 1), (24 * 16 * 16, 1024, ops.ACT_GLU, 1),
    (24 * 64 * 64, 128, ops.ACT_LRELU, 3), (24 * 32 * 32, 256, ops.ACT_LRELU, 3), (24 * 16 * 16, 512, ops.ACT_LRELU, 3),
    (24 * 8 * 8, 1024, ops.ACT_LRELU, 3), (24 * 4 * 4, 2048, ops.ACT_LRELU, 3), (24 * 32 * 32, 128, ops.ACT_LRELU, 1),
]
NAME = {ops.ACT_GLU: "glu", ops.ACT_LRELU: "lrelu", ops.ACT_NONE: "none"}
REPS = 10
for P, C, act, G in CASES:
    rows = P * G
    Co = C // 2 if act == ops.ACT_GLU else C
    in_b, out_b = rows * C * 2, rows * Co * 2
    ncopy = max(2, int(400e6 // (in_b + out_b)) + 1)
    ncopy = min(ncopy, 16)
    ys = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(ncopy)]
    ds = [torch.randn(rows, Co, device=dev).bfloat16() for _ in range(ncopy)]
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    st = torch.zeros(G * 2 * C, device=dev, dtype=torch.float64)
    ops.bn_stats(ys[0], st, groups=G)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    sts = [st.clone() for _ in range(REPS + 3)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    kw = dict(groups=G) if G > 1 else {}
    for i in range(3):
        out, mean, rstd = ops.bn_act_fwd(ys[i % ncopy], gamma, beta, act, stats=sts[REPS + i], **kw)
        ops.bn_act_bwd(ys[i % ncopy], ds[i % ncopy], mean, rstd, gamma, beta, act, dg, db, False, **kw)
    torch.cuda.synchronize()
    # the launches are captured into CUDA graphs so the host-side call overhead (~20 us / call) is not what is timed
    keep = []
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        for i in range(REPS):
            keep.append(ops.bn_act_fwd(ys[i % ncopy], gamma, beta, act, stats=sts[i], **kw))
    with torch.cuda.graph(gb):
        for i in range(REPS):
            keep.append(ops.bn_act_bwd(ys[i % ncopy], ds[i % ncopy], mean, rstd, gamma, beta, act, dg, db, False, **kw))
    gf.replay(); gb.replay()
    torch.cuda.synchronize()
    e[0].record()
    gf.replay()
    e[1].record()
    gb.replay()
    e[2].record()
    torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]) / REPS * 1e3, e[1].elapsed_time(e[2]) / REPS * 1e3
    fb = in_b + out_b                       # fwd: read x, write out
    bb = 2 * (in_b + out_b) + in_b          # bwd: reduce reads x + dout, apply reads x + dout, writes dx
    print(f"P={P:8d} x{G} C={C:5d} {NAME[act]:5s}  fwd {tf:7.1f} us {fb / tf / 1e3:7.0f} GB/s   "
          f"bwd {tb:7.1f} us {bb / tb / 1e3:7.0f} GB/s   ({fb / 1e6:.0f} / {bb / 1e6:.0f} MB)", flush=True)
    del ys, ds
