import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle.stackgan_oracle import Cfg, d_forward, is_param
from tests.parity_util import fp32_strict, make_d, rel
from sg2b200 import config
config.set_precision("fp32")
cfg = Cfg(); fp32_strict()
which, B = 0, 8
for case in ("cond", "uncond", "ximm"):
    net, sd = make_d(cfg, which, seed=2)
    for k in sd:
        if is_param(k): sd[k].requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    S = 64
    base = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
    c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    img, c = base.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    oimg, oc = base.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    (cond, uncond), x_imm = net(img * 1.0, c * 1.0)
    (ocond, ouncond), ox = d_forward(sd, oimg * 1.0, oc * 1.0, which, cfg, True)
    r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
    r3 = torch.randn(ox.shape, generator=g).cuda() * 0.01
    if case == "cond": (cond * r1).sum().backward(); (ocond * r1).sum().backward()
    if case == "uncond": (uncond * r2).sum().backward(); (ouncond * r2).sum().backward()
    if case == "ximm": (x_imm * r3).sum().backward(); (ox * r3).sum().backward()
    print(case, {k: f"{rel(p.grad, sd[k].grad):.2e}" for k, p in net.named_parameters() if sd[k].grad is not None and float(sd[k].grad.abs().max()) > 0},
          "d_img", f"{rel(img.grad, oimg.grad):.2e}", "d_c", f"{rel(c.grad, oc.grad):.2e}" if oc.grad is not None else None)
print("---- against a float64 run of the oracle")
net, sd = make_d(cfg, which, seed=2)
sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
for s_ in (sd, sd64):
    for k in s_:
        if is_param(k): s_[k].requires_grad_(True)
g = torch.Generator().manual_seed(5)
base = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).cuda()
c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
img, c = base.clone().requires_grad_(True), c0.clone().requires_grad_(True)
(cond, uncond), x_imm = net(img * 1.0, c * 1.0)
(oc_, ou_), ox_ = d_forward(sd, base.clone(), c0.clone(), which, cfg, True)
(dc_, du_), dx_ = d_forward(sd64, base.double(), c0.double(), which, cfg, True)
r2 = torch.randn(B, generator=g).cuda()
(uncond * r2).sum().backward(); (ou_ * r2).sum().backward(); (du_ * r2.double()).sum().backward()
for k, p in net.named_parameters():
    if sd64[k].grad is not None and float(sd64[k].grad.abs().max()) > 0:
        print(f"{k:28s} ours-vs-f64 {rel(p.grad, sd64[k].grad):.2e}   fp32oracle-vs-f64 {rel(sd[k].grad, sd64[k].grad):.2e}")
