"""In-graph time of each phase of the fused step, each captured alone in a CUDA graph (single stream):
G forward | D_i update (batched real/wrong/fake + Adam) | D_i part of the G step | G backward + Adam."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SG2_CONCURRENT"] = "0"
import torch
from sg2b200 import config, ops, trainer, utils
from sg2b200.nets import GradSink
_p, _st = ops._p, ops._st
B = 24
cfg = config.cfg
dev = torch.device("cuda:0")
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, dev)
tr = trainer.FusedTrainer(netG, netsD, cfg)
b = utils.synthetic_batch(cfg, B, seed=1, device=dev)
eps = torch.randn(B, cfg.GAN.EMBEDDING_DIM, device=dev)
for _ in range(2):
    tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=eps)
torch.cuda.synchronize()
state = {}


def ph_gfwd():
    ops.arena_reset(dev)
    state["g"] = tr.G.forward(b["z"], b["emb"], eps, True)


def ph_dupd(i):
    def f():
        fake, mu, logvar, Tg = state["g"]
        D, bucket = tr.Ds[i], tr.bD[i]
        bucket.grad.zero_()
        ready, fin = tr._layerwise(bucket, tr.lr_d)
        sink = GradSink(bucket.views, None, prezeroed=True, on_ready=ready)
        imgs3 = torch.cat((b["real"][i], b["wrong"][i], fake[i]), 0)
        probs = torch.empty(2, 3 * B, device=dev)
        _, _, _, T3 = D.forward(imgs3, mu.repeat(3, 1), True, probs[0], probs[1], groups=3)
        dprobs = tr._bce(probs.view(6, B), (1, 0, 0, 1, 1, 0), (1, 1, 1, 1, 1, 1), tr.losses[i:i + 1]).view(2, 3 * B)
        D.backward(T3, dprobs[0], dprobs[1], None, False, False, True, sink)
        sink.finish()
        fin()
    return f


def ph_dg(i):
    def f():
        fake, mu, logvar, Tg = state["g"]
        D = tr.Ds[i]
        probs = torch.empty(2, B, device=dev)
        _, _, x_imm, T = D.forward(fake[i], mu, True, probs[0], probs[1])
        dprobs = tr._bce(probs, (1, 1), (1, 1), tr.losses[3:4])
        ws = torch.empty(2 * B * B, device=dev)
        dx_imm = torch.empty_like(x_imm)
        ops._call("sg2_cal_loss", 3, _p(x_imm), _p(b["labels"]), B, x_imm.shape[1], _p(ws), _p(tr.losses[5:6]), _p(dx_imm), _st())
        state[("dimg", i)] = D.backward(T, dprobs[0], dprobs[1], dx_imm, True, True, False, None)
    return f


def ph_gbwd():
    fake, mu, logvar, Tg = state["g"]
    dimgs = [state[("dimg", i)][1] for i in range(3)]
    tr.bG.grad.zero_()
    ready, fin = tr._layerwise(tr.bG, tr.lr_g)
    sink = GradSink(tr.bG.views, None, prezeroed=True, on_ready=ready)
    tr.G.backward(Tg, dimgs, torch.zeros_like(mu), torch.zeros_like(mu), sink)
    sink.finish()
    fin()


def timed(name, fn):
    fn(); fn()
    torch.cuda.synchronize()
    n0 = ops.launches()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    nl = ops.launches() - n0
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:28s} {ms:7.3f} ms   {nl:4d} launches", flush=True)
    return ms


tot = timed("G forward", ph_gfwd)
for i in range(3):
    tot += timed(f"D{64 * 2 ** i} update (3B pass + Adam)", ph_dupd(i))
for i in range(3):
    tot += timed(f"D{64 * 2 ** i} G-step part", ph_dg(i))
tot += timed("G backward + Adam", ph_gbwd)
print(f"sum {tot:.3f} ms")
