"""Per-layer timing of every convolution (fprop / dgrad / wgrad) and BN/activation kernel of the 3-stage train step
at the benchmark shape, each timed alone with CUDA events. Prints executed TFLOP/s (convs) or GB/s (BN/act) per
launch and the time the layer contributes to one step (launch time x launches per step), sorted by contribution.

    gpurun -- 'python tools/layer_bench.py > gpurun_out/layers.log'
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sg2b200 import ops  # noqa: E402
from sg2b200.ops import ACT_GLU, ACT_LRELU, ACT_NONE, CONV3, CONV4S2, GEMM, UPCONV  # noqa: E402

B = int(os.environ.get("B", "24"))
REPS = int(os.environ.get("REPS", "20"))
dev = torch.device("cuda:0")

# name, kind, H(in), Cin, Cout, act, (n_fprop, n_dgrad, n_wgrad) per step
G = (1, 1, 1)
LAYERS = [
    ("G.up1", UPCONV, 4, 1024, 1024, ACT_GLU, G), ("G.up2", UPCONV, 8, 512, 512, ACT_GLU, G),
    ("G.up3", UPCONV, 16, 256, 256, ACT_GLU, G), ("G.up4", UPCONV, 32, 128, 128, ACT_GLU, G),
    ("G.head1", CONV3, 64, 64, 16, None, G),
    ("G.joint2", CONV3, 64, 192, 128, ACT_GLU, G), ("G.res2a", CONV3, 64, 64, 128, ACT_GLU, (2, 2, 2)),
    ("G.res2b", CONV3, 64, 64, 64, ACT_NONE, (2, 2, 2)), ("G.up_s2", UPCONV, 64, 64, 64, ACT_GLU, G),
    ("G.head2", CONV3, 128, 32, 16, None, G),
    ("G.joint3", CONV3, 128, 160, 64, ACT_GLU, G), ("G.res3a", CONV3, 128, 32, 64, ACT_GLU, (2, 2, 2)),
    ("G.res3b", CONV3, 128, 32, 32, ACT_NONE, (2, 2, 2)), ("G.up_s3", UPCONV, 128, 32, 32, ACT_GLU, G),
    ("G.head3", CONV3, 256, 16, 16, None, G),
]
D = (4, 4, 3)
for S in (64, 128, 256):
    LAYERS += [(f"D{S}.stem", GEMM, S // 2, 64, 64, None, (4, 1, 3)),
               (f"D{S}.s16_2", CONV4S2, S // 2, 64, 128, ACT_LRELU, D), (f"D{S}.s16_5", CONV4S2, S // 4, 128, 256, ACT_LRELU, D),
               (f"D{S}.s16_8", CONV4S2, S // 8, 256, 512, ACT_LRELU, D)]
    if S == 128:
        LAYERS += [("D128.s32", CONV4S2, 8, 512, 1024, ACT_LRELU, D), ("D128.s32_1", CONV3, 4, 1024, 512, ACT_LRELU, D)]
    if S == 256:
        LAYERS += [("D256.s32", CONV4S2, 16, 512, 1024, ACT_LRELU, D), ("D256.s64", CONV4S2, 8, 1024, 2048, ACT_LRELU, D),
                   ("D256.s64_1", CONV3, 4, 2048, 1024, ACT_LRELU, D), ("D256.s64_2", CONV3, 4, 1024, 512, ACT_LRELU, D)]
    LAYERS += [(f"D{S}.joint", CONV3, 4, 640, 512, ACT_LRELU, D)]


def timed(fn, reps=REPS):
    """us per launch, replaying a CUDA graph of `reps` launches (the eager path is CPU-launch-bound below ~30 us)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1000.0 * e0.elapsed_time(e1) / reps      # us


def main():
    only = os.environ.get("ONLY")
    rows = []
    for name, kind, H, Cin, Cout, act, cnt in LAYERS:
        if only and only not in name:
            continue
        if kind == GEMM:                       # stem: (1,1,P,64) x (64 -> 64)
            P = B * H * H
            x = torch.randn(1, 1, P, Cin, device=dev).bfloat16()
            Bx, Hx, Wx = 1, 1, P
        else:
            x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
            Bx, Hx, Wx = B, H, H
        s1, s2 = ops.pack_shapes(kind, Cout, Cin)
        wpk = (torch.randn(s1, device=dev) * 0.02).bfloat16()
        wpkT = (torch.randn(s2, device=dev) * 0.02).bfloat16()
        Ho, Wo = ops._out_hw(kind, Hx, Wx)
        dy = torch.randn(Bx, Ho, Wo, Cout, device=dev).bfloat16()
        dwpk = torch.zeros(Cout, ops.JOBS[kind], Cin, device=dev, dtype=torch.float32)
        fl = ops._conv_flops(kind, Bx, Hx, Wx, Cin, Cout)
        has_bn = act is not None
        st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64)
        t_f = timed(lambda: ops.conv_fprop(kind, x, wpk, Cout, stats=st if has_bn else None))
        t_d = timed(lambda: ops.conv_dgrad(kind, dy, wpkT, Bx, Hx, Wx, Cin))
        t_w = timed(lambda: ops.conv_wgrad(kind, x, dy, dwpk))
        rows.append((name, "fprop", t_f, fl / t_f / 1e6, cnt[0]))
        rows.append((name, "dgrad", t_d, fl / t_d / 1e6, cnt[1]))
        rows.append((name, "wgrad", t_w, fl / t_w / 1e6, cnt[2]))
        if has_bn:
            y = dy
            gamma, beta = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
            Ca = Cout // 2 if act == ACT_GLU else Cout
            dout = torch.randn(Bx, Ho, Wo, Ca, device=dev).bfloat16()
            st.zero_()
            ops.bn_stats(y.view(-1, Cout), st)
            res = ops.bn_act_fwd(y, gamma, beta, act, stats=st)
            mean, rstd = res[1], res[2]
            st2 = st.clone()

            def f_fwd():
                st.copy_(st2)
                ops.bn_act_fwd(y, gamma, beta, act, stats=st)
            dg, db = torch.zeros(Cout, device=dev), torch.zeros(Cout, device=dev)
            t_bf = timed(f_fwd)
            t_bb = timed(lambda: ops.bn_act_bwd(y, dout, mean, rstd, gamma, beta, act, dg, db, False))
            nel = y.numel()
            byf = nel * 2 + dout.numel() * 2
            byb = 2 * (nel * 2 + dout.numel() * 2) + nel * 2
            rows.append((name, "bn_fwd", t_bf, byf / t_bf / 1e3, cnt[0]))
            rows.append((name, "bn_bwd", t_bb, byb / t_bb / 1e3, cnt[1]))
    tot = sum(r[2] * r[4] for r in rows)
    print(f"B={B}: sum over layers of launch time x launches/step = {tot/1000:.3f} ms")
    print(f"{'layer':12s} {'op':7s} {'us':>9s} {'TF/s|GB/s':>10s} {'n/step':>6s} {'us/step':>9s} {'share':>6s}")
    for r in sorted(rows, key=lambda r: -r[2] * r[4]):
        print(f"{r[0]:12s} {r[1]:7s} {r[2]:9.1f} {r[3]:10.1f} {r[4]:6d} {r[2]*r[4]:9.1f} {100*r[2]*r[4]/tot:5.1f}%")
    for op in ("fprop", "dgrad", "wgrad", "bn_fwd", "bn_bwd"):
        print(f"total {op:7s} {sum(r[2]*r[4] for r in rows if r[1]==op)/1000:8.3f} ms")


if __name__ == "__main__":
    main()
