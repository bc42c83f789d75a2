"""Replays the rotating-batch run to the step before the first non-finite gradient (tools/scratch/find_nan.py: step 189),
then runs that step with every sg2b200.ops call checked: prints the first ops whose returned tensors are non-finite."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import config, ops, trainer, utils
cfg = config.cfg
dev = torch.device("cuda:0")
BAD = int(os.environ.get("BAD_STEP", "189"))
B = 24
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, dev)
tr = trainer.FusedTrainer(netG, netsD, cfg)
batches = [utils.synthetic_batch(cfg, B, seed=1000 + i, device=dev, n_classes=200) for i in range(3)]
found = []


def tensors(o):
    if isinstance(o, torch.Tensor):
        yield o
    elif isinstance(o, (tuple, list)):
        for x in o:
            yield from tensors(x)


def wrap(name, fn):
    def w(*a, **k):
        out = fn(*a, **k)
        if len(found) < 12:
            torch.cuda.synchronize()
            ins = [t for t in tensors(a) if t.is_floating_point()]
            bad_in = [tuple(t.shape) for t in ins if not torch.isfinite(t).all()]
            for t in tensors(out):
                if t.is_floating_point() and not torch.isfinite(t).all():
                    mx = max((float(x.float().abs().max()) for x in ins if torch.isfinite(x).all() and x.numel()), default=0.0)
                    found.append(name)
                    print(f"NON-FINITE out of ops.{name}: out {tuple(t.shape)} {t.dtype} bad={int((~torch.isfinite(t)).sum())} "
                          f"| non-finite inputs: {bad_in} | max |finite input| {mx:.3e} | arg shapes "
                          f"{[tuple(x.shape) for x in ins][:6]} scalars {[x for x in a if isinstance(x, (int, float))][:8]}", flush=True)
                    break
        return out
    return w


for s in range(BAD + 1):
    b = batches[s % 3]
    eps = torch.randn(B, cfg.GAN.EMBEDDING_DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(s))
    if s == BAD:
        for n in dir(ops):
            f = getattr(ops, n)
            if isinstance(f, types.FunctionType) and not n.startswith("_") and n not in ("launches", "arena_reset", "bn_stats32"):
                setattr(ops, n, wrap(n, f))
    lo = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=eps)
torch.cuda.synchronize()
print("losses", lo.tolist(), "found", found)
