"""stem_im2col / stem_col2im on the three image scales (graph replays; L2-warm and behind a 256 MB fill)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sg2b200 import ops
from layer_bench import timed
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "24"))
flush = torch.empty(64 << 20, device=dev)
f0 = timed(lambda: flush.zero_())
for S in (64, 128, 256):
    img = torch.randn(B, 3, S, S, device=dev)
    col = torch.empty(B * (S // 2) ** 2, 64, device=dev, dtype=torch.bfloat16)
    dcol = torch.randn(B * (S // 2) ** 2, 64, device=dev).bfloat16()
    a = timed(lambda: ops.stem_im2col(img, out=col)); a2 = timed(lambda: (flush.zero_(), ops.stem_im2col(img, out=col))) - f0
    b = timed(lambda: ops.stem_col2im(dcol, B, S)); b2 = timed(lambda: (flush.zero_(), ops.stem_col2im(dcol, B, S))) - f0
    mb = (img.numel() * 4 + col.numel() * 2) / 1e6
    print(f"S={S} B={B}: im2col {a:6.1f} us warm {a2:6.1f} us cold ({mb / a2 / 1e3:.2f} TB/s)   col2im {b:6.1f} us warm {b2:6.1f} us cold")
