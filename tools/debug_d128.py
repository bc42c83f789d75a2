"""Debug: 3-stage fused training for N steps; prints errD per step and parameter / buffer magnitudes of D128."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import config, trainer, utils
N = int(os.environ.get("N", "160"))
cfg = config.cfg
cfg.TREE.BRANCH_NUM = 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, dev)
tr = trainer.FusedTrainer(netG, netsD, cfg)
for s in range(N):
    b = utils.synthetic_batch(cfg, 24, seed=1000 + s, device=dev, n_classes=6)
    eps = torch.randn(24, cfg.GAN.EMBEDDING_DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(s))
    lo = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=eps).cpu().tolist()
    if s % 5 == 0 or lo[1] > 4:
        print(s, [round(v, 3) for v in lo], flush=True)
for k, v in netsD[1].state_dict().items():
    v = v.float()
    print(f"{k:40s} absmax {float(v.abs().max()):10.4f} mean {float(v.mean()):10.4f} finite {bool(torch.isfinite(v).all())}")
