"""GPU parity of the whole train step: sg2b200.trainer.FusedTrainer vs the oracle's OracleTrainer (fp32, TF32 off)
on identical weights and inputs — every loss of the step, the updated weights, the EMA shadow and the BN running
statistics — plus the drop-in module API driven through autograd, and a multi-step loss curve."""
import pytest
import torch

from oracle.stackgan_oracle import Cfg, OracleTrainer, param_keys
from tests.parity_util import fp32_strict, rel, set_cfg

pytestmark = pytest.mark.gpu


def _setup(branches, B, seed=0):
    from sg2b200 import trainer, utils
    fp32_strict()
    ocfg = Cfg(BRANCH_NUM=branches)
    cfg = set_cfg(ocfg)
    torch.manual_seed(seed)
    netG, netsD = utils.build_networks(cfg, "cuda")
    orc = OracleTrainer(ocfg, {k: v.detach().clone() for k, v in netG.state_dict().items()},
                        [{k: v.detach().clone() for k, v in d.state_dict().items()} for d in netsD], device="cuda")
    tr = trainer.FusedTrainer(netG, netsD, cfg)
    return cfg, ocfg, netG, netsD, tr, orc


def _batch(cfg, B, seed):
    from sg2b200 import utils
    b = utils.synthetic_batch(cfg, B, seed=seed, device="cuda", n_classes=3)
    b["eps"] = torch.randn(B, cfg.GAN.EMBEDDING_DIM, generator=torch.Generator().manual_seed(seed + 1)).cuda()
    return b


def _ostep(orc, b):
    return orc.step(dict(z=b["z"], emb=b["emb"], eps=b["eps"], real=b["real"], wrong=b["wrong"],
                         labels=b["labels"].tolist()))


def _flat(d, prefix):
    return torch.cat([v.detach().double().flatten() for k, v in d.items() if k.startswith(prefix)])


@pytest.mark.parametrize("branches,B", [(1, 8), (3, 24)])   # BASELINE.json configs[0] (shape) and configs[1]
def test_fused_step_losses_and_gradients(branches, B):
    """One whole step (lr = 0: a comparison of gradients; the G step sees identical D weights in every arm) against the
    FLOAT64 oracle: all losses within 5e-3 (measured <= 7e-4), and EVERY parameter gradient of G and the discriminators,
    per network, as close to the truth as an ideal implementation of bf16 storage (the bf16-emulating oracle; factor
    YARD = 1.3, measured 0.99 - 1.04; see tests/test_gpu_nets.py for the method and profiles/r02_parity.md for numbers:
    the bf16 quantisation gap itself is 6e-2 ... 2e-1 on these gradients, cosine 0.977 - 0.998)."""
    from oracle.stackgan_oracle import OracleTrainer, emulate_bf16
    from tests.parity_util import (bucket_grads, build_trainer_and_oracles, f64_state, loss_vector, oracle_grads,
                                   oracle_step, train_batch)
    cfg, ocfg, netG, netsD, tr, (o32, oq) = build_trainer_and_oracles(branches, seed=0, n_oracles=2, lr=0.0)
    ot = OracleTrainer(ocfg, f64_state({k: v.detach() for k, v in o32.g.items()}),
                       [f64_state({k: v.detach() for k, v in d.items()}) for d in o32.ds], device="cuda")
    b = train_batch(cfg, B, 11)
    losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu().tolist()
    rt = oracle_step(ot, b, keep_grads=True)
    with emulate_bf16():
        rq = oracle_step(oq, b, keep_grads=True)
    for name, a, r in zip([f"errD{i}" for i in range(branches)] + ["errG_total", "kl", "cal"], losses, loss_vector(rt)):
        assert abs(a - r) <= 5e-3 * abs(r) + 2e-4, (name, a, r)
    ours, emu, truth = bucket_grads(tr), oracle_grads(rq), oracle_grads(rt)
    for n in ["G"] + [f"D{i}" for i in range(branches)]:
        e, eq = rel(_flat(ours, n + "."), _flat(truth, n + ".")), rel(_flat(emu, n + "."), _flat(truth, n + "."))
        assert e <= 1.3 * eq + 2e-3, (n, e, eq)
    sd = netG.state_dict()
    for k in ot.g:
        if "running" in k:
            assert rel(sd[k].float(), ot.g[k].float()) < 2e-2, k
        if "num_batches" in k:
            assert int(sd[k]) == int(ot.g[k]) == 1
    for d, osd in zip(netsD, ot.ds):
        sdd = d.state_dict()
        assert all(rel(sdd[k].float(), osd[k].float()) < 2e-2 for k in osd if "running" in k)
        assert all(int(sdd[k]) == int(osd[k]) == 4 for k in osd if "num_batches" in k)   # 3 D-step + 1 G-step passes


@pytest.mark.parametrize("branches,B", [(1, 8), (3, 6), (3, 24)])   # (3, 24) = BASELINE.json's configs[1]
def test_fused_step_matches_oracle(branches, B):
    """Losses within 2e-2 relative (bf16 activations vs fp32 reference). Adam's first step moves every weight by
    lr * g/(|g| + eps) ~ +-lr, so after the update two implementations can differ by at most 2 lr per element (a
    gradient whose sign differs); the mean difference must stay well below that (most signs agree)."""
    cfg, ocfg, netG, netsD, tr, orc = _setup(branches, B)
    b = _batch(cfg, B, 11)
    losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu()
    o = _ostep(orc, b)
    ref = [float(e) for e in o["errD"]] + [float(o["errG_total"]), float(o["kl"]), float(o["cal"])]
    for name, a, r in zip([f"errD{i}" for i in range(branches)] + ["errG_total", "kl", "cal"], losses.tolist(), ref):
        assert abs(a - r) <= 2e-2 * abs(r) + 2e-4, (name, a, r)
    lr = 2e-4

    def close_after_adam(a, r, k):
        d = (a.detach() - r.detach()).abs()
        assert float(d.max()) <= 2.05 * lr, (k, float(d.max()))
        if d.numel() >= 4096:       # tiny tensors (BN biases start at 0 with near-zero gradients): sign is noise
            assert float(d.mean()) <= 0.35 * lr, (k, float(d.mean()))

    sd = netG.state_dict()
    for k in param_keys(orc.g):
        close_after_adam(sd[k], orc.g[k], k)
    for k in orc.g:
        if "running" in k:
            assert rel(sd[k].float(), orc.g[k].float()) < 2e-2, k
        if "num_batches" in k:
            assert int(sd[k]) == int(orc.g[k]) == 1
    for d, osd in zip(netsD, orc.ds):
        sdd = d.state_dict()
        for k in param_keys(osd):
            close_after_adam(sdd[k], osd[k], k)
        assert all(int(sdd[k]) == int(osd[k]) == 4 for k in osd if "num_batches" in k)   # 3 D-step + 1 G-step passes
    for a, r in zip(tr.bG.ema_params(), orc.avg_g):
        assert float((a - r).abs().max()) <= 2.05e-3 * lr + 2e-7        # 0.001 * (2 lr) + fp32 rounding of O(1) values


def test_loss_curve_tracks_oracle():
    """8 consecutive steps (fresh z / eps / data each step). Adam moves every weight by ~lr per step whatever the
    gradient's size, so sign differences of near-zero gradients (bf16 vs fp32) make the two GAN trajectories drift
    apart slowly: the first step must agree to 2 %, the curve must stay within 10 % over the 8 steps."""
    cfg, ocfg, netG, netsD, tr, orc = _setup(1, 8, seed=3)
    for s in range(8):
        b = _batch(cfg, 8, 100 + s)
        losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu().tolist()
        o = _ostep(orc, b)
        ref = [float(o["errD"][0]), float(o["errG_total"]), float(o["kl"])]
        tol = 2e-2 if s == 0 else 1e-1
        for a, r in zip(losses[:3], ref):
            assert abs(a - r) <= tol * abs(r) + 1e-3, (s, losses, ref)


def test_module_api_autograd_step_matches_fused():
    """The drop-in path (G_NET / D_NET called like the reference's train_Dnet / train_Gnet do, losses by torch,
    gradients by autograd) produces the same gradients as the fused trainer's hand-scheduled backward.
    Two runs of the SAME kernels are not bit-identical: split-K layers accumulate with fp32 atomics, a handful of
    bf16 roundings flip, and small-batch BatchNorm + LeakyReLU masks amplify that to ~1e-2 in D's gradients
    (tools/debug_determinism.py). Hence cosine >= 0.995 and l2 <= 8e-2 rather than equality."""
    from oracle.stackgan_oracle import bce, class_aware_loss, kl_loss
    from sg2b200 import utils
    cfg, ocfg, netG, netsD, tr, orc = _setup(1, 8, seed=5)
    b = _batch(cfg, 8, 21)
    ones, zeros = torch.ones(8, device="cuda"), torch.zeros(8, device="cuda")
    # reference-style D step through the module API
    for p in netsD[0].parameters():
        p.grad = None
    fake, mu, logvar = netG(b["z"], b["emb"], eps=b["eps"])
    rl, _ = netsD[0](b["real"][0], mu.detach())
    wl, _ = netsD[0](b["wrong"][0], mu.detach())
    fl, _ = netsD[0](fake[0].detach(), mu.detach())
    errD = (bce(rl[0], ones) + bce(rl[1], ones)) + (bce(wl[0], zeros) + bce(wl[1], ones)) + (bce(fl[0], zeros) + bce(fl[1], zeros))
    errD.backward()
    gD = {k: p.grad.clone() for k, p in netsD[0].named_parameters()}
    # the same D step inside the fused trainer (lr = 0 so that weights stay put), then compare its flat gradient bucket
    tr.lr_d = tr.lr_g = 0.0
    losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu()
    assert abs(float(losses[0]) - float(errD)) <= 1e-2 * abs(float(errD))
    import torch.nn.functional as F
    for k, p in netsD[0].named_parameters():
        a, r = tr.bD[0].views[p].flatten().double(), gD[k].flatten().double()
        assert float(F.cosine_similarity(a, r, dim=0)) > 0.995, k
        assert rel(a, r) < 8e-2, k
