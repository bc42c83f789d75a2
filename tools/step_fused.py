"""Two fused train steps at the benchmark shape (for ncu launch lists / ncu --set full captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import config, ops, trainer, utils

B = int(os.environ.get("B", "24"))
STEPS = int(os.environ.get("STEPS", "2"))
cfg = config.cfg
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, "cuda")
tr = trainer.FusedTrainer(netG, netsD, cfg)
b = utils.synthetic_batch(cfg, B, seed=1, device="cuda")
for s in range(STEPS):
    n0 = ops.launches()
    losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"])
    torch.cuda.synchronize()
    print("step", s, "launches", ops.launches() - n0, "losses", [round(v, 4) for v in losses.tolist()])
