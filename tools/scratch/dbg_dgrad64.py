"""dgrad of conv4x4-s2 64->128 @128x128 (D256's first BN layer): with / without the epilogue operand, B = 24 / 72."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sg2b200 import ops
import layer_bench as lb
dev = torch.device("cuda:0")
for B in (24, 72):
    for (kind, H, Cin, Cout) in ((ops.CONV4S2, 128, 64, 128), (ops.CONV4S2, 64, 128, 256)):
        s1, s2 = ops.pack_shapes(kind, Cout, Cin)
        wpkT = (torch.randn(s2, device=dev) * 0.02).bfloat16()
        wpk = (torch.randn(s1, device=dev) * 0.02).bfloat16()
        dy = torch.randn(B, H // 2, H // 2, Cout, device=dev).bfloat16()
        x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
        src = torch.randn(B, H, H, Cin, device=dev).bfloat16()
        fl = 2.0 * B * (H // 2) ** 2 * 16 * Cin * Cout
        for name, fn in (("dgrad", lambda: ops.conv_dgrad(kind, dy, wpkT, B, H, H, Cin)),
                         ("dgrad+mask", lambda: ops.conv_dgrad(kind, dy, wpkT, B, H, H, Cin, epi=(src, ops.EPI_LRELU_MASK))),
                         ("dgrad+add", lambda: ops.conv_dgrad(kind, dy, wpkT, B, H, H, Cin, epi=(src, ops.EPI_ADD))),
                         ("fprop", lambda: ops.conv_fprop(kind, x, wpk, Cout))):
            us = lb.timed(fn)
            print(f"B={B} {H}x{H} {Cin}->{Cout} {name:11s} {us:7.1f} us {fl / us / 1e6:7.1f} TF/s", flush=True)
