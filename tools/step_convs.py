"""Per-launch timing of every conv launch of one real train step (recorded, then replayed one by one in CUDA graphs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import config, ops, trainer, utils
B = int(os.environ.get("B", "24"))
cfg = config.cfg
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, "cuda")
tr = trainer.FusedTrainer(netG, netsD, cfg)
b = utils.synthetic_batch(cfg, B, seed=1, device="cuda")
for _ in range(2):
    tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"])
torch.cuda.synchronize()
ops.profile_begin()
tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"])
torch.cuda.synchronize()
calls, keep = ops.profile_end()
KIND = {0: "conv3", 1: "upconv", 2: "conv4s2", 3: "gemm"}
rows = []
for name, a, fl in calls:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5):
            ops.replay_call(name, a)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = 1000 * e0.elapsed_time(e1) / 5
    op = name.replace("sg2_conv_", "")
    if op == "wgrad":
        kind, Bn, H, W, Ci, Co, sk = a[0], a[4], a[5], a[6], a[7], a[8], a[9]
    elif op == "fprop":
        kind, Bn, H, W, Ci, Co, sk = a[0], a[5], a[6], a[7], a[8], a[9], a[10]
    else:
        kind, Bn, H, W, Ci, Co, sk = a[0], a[5], a[6], a[7], a[8], a[9], a[10]
    rows.append((us, op, KIND[kind], Bn, H, W, Ci, Co, sk, fl / us / 1e6))
tot = sum(r[0] for r in rows)
print(f"{len(rows)} conv launches, {tot/1000:.3f} ms, {sum(c[2] for c in calls)/tot/1e6:.1f} TF/s")
for r in sorted(rows, key=lambda r: -r[0]):
    print(f"{r[0]:8.1f} us {100*r[0]/tot:5.1f}%  {r[1]:5s} {r[2]:8s} B={r[3]:<5d} {r[4]:4d}x{r[5]:<6d} {r[6]:5d}->{r[7]:<5d} splitk={r[8]:<3d} {r[9]:7.1f} TF/s")
