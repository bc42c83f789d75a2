"""Rotating-batch run (bench.py's data pattern) until the first non-finite value: which step, which bucket, which
parameters' gradients / weights. Run twice in one process: same step both times = deterministic numerics, not a race."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import config, trainer, utils
cfg = config.cfg
dev = torch.device("cuda:0")
STEPS = int(os.environ.get("STEPS", "260"))
B = int(os.environ.get("B", "24"))


def run(tag):
    torch.manual_seed(0)
    netG, netsD = utils.build_networks(cfg, dev)
    tr = trainer.FusedTrainer(netG, netsD, cfg)
    names = [{p: n for n, p in net.named_parameters()} for net in [netG] + list(netsD)]
    batches = [utils.synthetic_batch(cfg, B, seed=1000 + i, device=dev, n_classes=200) for i in range(3)]
    for s in range(STEPS):
        b = batches[s % 3]
        eps = torch.randn(B, cfg.GAN.EMBEDDING_DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(s))
        lo = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=eps)
        torch.cuda.synchronize()
        bad = False
        for bi, bk in enumerate([tr.bG] + tr.bD):
            for what in ("grad", "flat"):
                t = getattr(bk, what)
                if not torch.isfinite(t).all():
                    bad = True
                    offenders = []
                    for p, (o, n) in bk.range_of.items():
                        sl = t[o:o + n]
                        if not torch.isfinite(sl).all():
                            offenders.append(f"{names[bi][p]}({int((~torch.isfinite(sl)).sum())}/{n})")
                    print(f"[{tag}] step {s}: bucket {bi} ({'G' if bi == 0 else 'D%d' % (bi - 1)}) {what} non-finite in: {offenders[:12]}")
        if bad or s % 50 == 0:
            print(f"[{tag}] step {s} losses {[round(float(v), 4) for v in lo.tolist()]}", flush=True)
        if bad:
            gm = tr.bG.grad.abs()
            print(f"[{tag}] max |G grad| finite part {float(gm[torch.isfinite(gm)].max()):.3e}")
            return s
    print(f"[{tag}] finite through {STEPS} steps")
    return None


a = run("run1")
b = run("run2")
print("first non-finite step:", a, b)
