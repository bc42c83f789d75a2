"""Run the module-API passes of one train step at the production shape (for ncu launch lists): G fwd+bwd, D0-2 fwd+bwd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.stackgan_oracle import Cfg
from tests.parity_util import make_d, make_g

B = int(os.environ.get("B", "24"))
ITERS = int(os.environ.get("ITERS", "2"))
cfg = Cfg()
net, _ = make_g(cfg, seed=1)
ds = [make_d(cfg, i)[0] for i in range(3)]
z = torch.randn(B, cfg.Z_DIM).cuda(); emb = torch.randn(B, cfg.TEXT_DIM).cuda()
for it in range(ITERS):
    imgs, mu, lv = net(z, emb)
    (sum(i.mean() for i in imgs) + mu.mean() + lv.mean()).backward()
    for i, d in enumerate(ds):
        (cnd, unc), x = d(imgs[i].detach(), mu.detach())
        (cnd.mean() + unc.mean()).backward()
    torch.cuda.synchronize()
from sg2b200 import ops
print("launches per iteration:", ops.launches() // ITERS)
