import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.stackgan_oracle import Cfg, OracleTrainer, bce, param_keys
from tests.parity_util import fp32_strict, rel, set_cfg
from tests.test_gpu_train_step import _setup, _batch, _ostep
from sg2b200.nets import GradSink

cfg, ocfg, netG, netsD, tr, orc = _setup(1, 8, seed=5)
b = _batch(cfg, 8, 21)
D = netsD[0]; eng = D.engine()
ones, zeros = torch.ones(8, device="cuda"), torch.zeros(8, device="cuda")
fake, mu, logvar = netG(b["z"], b["emb"], eps=b["eps"])
mu_d = mu.detach()
# per-pass comparison: module API vs direct engine with external sink
for name, img, tc, tu in (("real", b["real"][0], ones, ones), ("wrong", b["wrong"][0], zeros, ones), ("fake", fake[0].detach(), zeros, zeros)):
    for p in D.parameters(): p.grad = None
    (c1, u1), _ = D(img, mu_d)
    (bce(c1, tc) + bce(u1, tu)).backward()
    g_mod = {k: p.grad.clone() for k, p in D.named_parameters()}
    # direct engine
    probs = torch.empty(2, 8, device="cuda")
    _, _, _, T = eng.forward(img.contiguous(), mu_d, True, probs[0], probs[1])
    print(name, "fwd equal:", float((probs[0]-c1).abs().max()), float((probs[1]-u1).abs().max()))
    dpr = tr._bce(probs, (float(tc[0]), float(tu[0])), (1, 1), torch.zeros(1, device="cuda"))
    # reference dprob
    pc = c1.detach().clone().requires_grad_(True); pu = u1.detach().clone().requires_grad_(True)
    gg = torch.autograd.grad(bce(pc, tc) + bce(pu, tu), (pc, pu))
    print("   dprob diff", float((dpr[0]-gg[0]).abs().max()), float((dpr[1]-gg[1]).abs().max()))
    sink = GradSink()
    eng.backward(T, dpr[0], dpr[1], None, False, False, True, sink)
    g_dir = sink.finish()
    worst = max(((k, rel(g_dir[p], g_mod[k])) for k, p in D.named_parameters()), key=lambda t: t[1])
    print("   module vs direct-engine worst:", worst)
    # with a views-backed sink (as the fused trainer uses)
    sink = GradSink(tr.bD[0].views)
    eng.backward(T, dpr[0], dpr[1], None, False, False, True, sink)
    sink.finish()
    worst = max(((k, rel(tr.bD[0].views[p], g_mod[k])) for k, p in D.named_parameters()), key=lambda t: t[1])
    print("   module vs views-sink worst:", worst)
