"""Run a D engine forward+backward twice on identical inputs, logging every ops.* output; report divergences."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.test_gpu_train_step import _setup, _batch
from sg2b200 import ops
from sg2b200.nets import GradSink

LOG = []
def wrap(name, fn):
    def w(*a, **k):
        out = fn(*a, **k)
        outs = out if isinstance(out, tuple) else (out,)
        for i, o in enumerate(outs):
            if torch.is_tensor(o):
                LOG.append((f"{name}[{i}] {tuple(o.shape)}", o.detach().float().clone()))
        if name in ("conv_wgrad",):
            LOG.append((f"{name} dwpk {tuple(a[3].shape)}", a[3].detach().clone()))
        if name == "logits_bwd":
            LOG.append((f"{name} dx", a[4].detach().float().clone()))
        if name == "concat_c_bwd":
            LOG.append((f"{name} dc", a[2].detach().clone()))
        return out
    return w
for n in ["conv_fprop", "conv_dgrad", "conv_wgrad", "f32_to_bf16_stats", "bn_act_fwd", "bn_act_bwd", "lrelu_bwd", "add_bf16",
          "f32_to_bf16", "concat_c", "concat_c_bwd", "stem_im2col", "logits_fwd", "logits_bwd", "nhwc_to_nchw_f32"]:
    setattr(ops, n, wrap(n, getattr(ops, n)))

which = int(os.environ.get("WHICH", "0"))
cfg, ocfg, netG, netsD, tr, orc = _setup(which + 1, 8, seed=5)
b = _batch(cfg, 8, 21)
eng = netsD[which].engine()
mu = torch.randn(8, 128, device="cuda")
runs = []
for r in range(2):
    LOG.clear()
    probs = torch.empty(2, 8, device="cuda")
    _, _, _, T = eng.forward(b["wrong"][which], mu, True, probs[0], probs[1])
    dpr = torch.randn(2, 8, generator=torch.Generator().manual_seed(1)).cuda()
    sink = GradSink()
    eng.backward(T, dpr[0], dpr[1], None, True, True, True, sink)
    sink.finish()
    torch.cuda.synchronize()
    runs.append(list(LOG))
print("entries", len(runs[0]), len(runs[1]))
for (n0, t0), (n1, t1) in zip(*runs):
    d = (t0 - t1).abs().max().item()
    r = ((t0 - t1).norm() / (t1.norm() + 1e-30)).item()
    nd = int((t0 != t1).sum())
    flag = "  <<<" if r > 1e-5 else ""
    print(f"{n0:48s} maxabs {d:.3e} rel {r:.3e} ndiff {nd}{flag}")
