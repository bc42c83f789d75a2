"""GPU parity: sg2b200 G_NET / D_NET* (CUDA kernels through the C ABI) vs the oracle on identical weights + inputs.

Metric: per-tensor relative error ||a-b|| / ||b|| against the fp32 oracle (TF32 off). The CUDA path stores every
activation in bf16 (fp32 accumulation), so the error grows with depth: measured on B200 (profiles/r01_parity.md)
img64 0.8e-2, img128 1.2e-2, img256 1.9e-2 after 8 / 14 / 20 conv+BN layers — the north-star 1e-2 holds per layer
(tests/test_gpu_kernels.py: <= 5e-3) and for the first stage, not end-to-end through 20 bf16 layers.
Gradients through LeakyReLU additionally see mask flips of pre-activations that lie within bf16 rounding of zero
(each flip changes a local derivative from 1 to 0.2), which dominates the l2 error of D's gradients; they are
therefore checked by cosine similarity as well. Tolerances below = measured value + margin, stated per check."""
import pytest
import torch

from oracle.stackgan_oracle import Cfg, d_forward, g_forward, is_param
from tests.parity_util import fp32_strict, make_d, make_g, rel, report

pytestmark = pytest.mark.gpu
FWD_TOL = 2.5e-2       # G images after up to 20 bf16 layers (measured <= 1.9e-2); mu / logvar are fp32 (1e-6)
D_FWD_TOL = 2e-2       # D logits / x_immediate (measured <= 1.2e-2)
G_GRAD_TOL = 7e-2      # GLU is smooth: measured 1-5e-2, cosine >= 0.999
D_GRAD_COS = 0.98      # LeakyReLU mask flips: measured cosine 0.986-0.9999, l2 5-15e-2


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def _g_case(cfg, B, training=True):
    fp32_strict()
    net, sd = make_g(cfg, seed=1)
    net.train(training)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, cfg.Z_DIM, generator=g).cuda()
    emb = torch.randn(B, cfg.TEXT_DIM, generator=g).cuda()
    eps = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    return net, sd, z, emb, eps, g


@pytest.mark.parametrize("branches,B", [(1, 8), (3, 4)])
def test_g_forward_backward(branches, B):
    cfg = Cfg(BRANCH_NUM=branches)
    net, sd, z, emb, eps, g = _g_case(cfg, B)
    for k in sd:
        if is_param(k):
            sd[k].requires_grad_(True)
    imgs, mu, logvar = net(z, emb, eps=eps)
    oimgs, omu, ologvar = g_forward(sd, z, emb, eps, cfg, True)
    pairs = [(f"img{i}", rel(a, b)) for i, (a, b) in enumerate(zip(imgs, oimgs))]
    pairs += [("mu", rel(mu, omu)), ("logvar", rel(logvar, ologvar))]
    ok, msg = report(pairs, FWD_TOL)
    assert ok, msg
    # BN running statistics updated like nn.BatchNorm (momentum 0.1, unbiased var)
    st = net.state_dict()
    pairs = [(k, rel(st[k].float(), sd[k].float())) for k in sd if "running" in k]
    ok, msg = report(pairs, FWD_TOL)
    assert ok, msg
    assert all(int(st[k]) == int(sd[k]) for k in sd if "num_batches" in k)
    # backward: random cotangents on every output
    rs = [torch.randn(i.shape, generator=g).cuda() for i in oimgs]
    rmu, rlv = torch.randn(mu.shape, generator=g).cuda(), torch.randn(mu.shape, generator=g).cuda()
    loss = sum((a * r).sum() for a, r in zip(imgs, rs)) + (mu * rmu).sum() + (logvar * rlv).sum()
    oloss = sum((a * r).sum() for a, r in zip(oimgs, rs)) + (omu * rmu).sum() + (ologvar * rlv).sum()
    loss.backward()
    oloss.backward()
    pairs = [(k, rel(p.grad, sd[k].grad)) for k, p in net.named_parameters()]
    ok, msg = report(pairs, G_GRAD_TOL)
    assert ok, msg
    worst = min(cos(p.grad, sd[k].grad) for k, p in net.named_parameters())
    assert worst > 0.998, worst


def test_g_eval_mode_uses_running_stats():
    cfg = Cfg(BRANCH_NUM=3)
    net, sd, z, emb, eps, g = _g_case(cfg, 4, training=False)
    with torch.no_grad():
        imgs, mu, logvar = net(z, emb, eps=eps)
        oimgs, _, _ = g_forward(sd, z, emb, eps, cfg, False)
    ok, msg = report([(f"img{i}", rel(a, b)) for i, (a, b) in enumerate(zip(imgs, oimgs))], FWD_TOL)
    assert ok, msg


def test_g_eval_synthesis_batch_256_equals_its_chunks():
    """netG.eval() no_grad synthesis at batch 256 (the reference's evaluate(), trainer.py:681-803).
    Eval-mode samples are independent, so the full-size batch must equal the concatenation of four 64-sample calls
    (different tile / split-K schedules and atomic summation orders flip individual bf16 roundings, which then propagate
    through up to 20 layers: measured 6e-3 at 256x256, bound 1.5e-2) — a size-independent check at a size the fp32
    oracle is not run at; the first 16 samples are also compared with the oracle."""
    cfg = Cfg(BRANCH_NUM=3)
    net, sd, _, _, _, g = _g_case(cfg, 4, training=False)
    gen = torch.Generator().manual_seed(9)
    B = 256
    z = torch.randn(B, cfg.Z_DIM, generator=gen).cuda()
    emb = torch.randn(B, cfg.TEXT_DIM, generator=gen).cuda()
    eps = torch.randn(B, cfg.EMBEDDING_DIM, generator=gen).cuda()
    with torch.no_grad():
        full, mu, _ = net(z, emb, eps=eps)
        parts = [net(z[i:i + 64], emb[i:i + 64], eps=eps[i:i + 64])[0] for i in range(0, B, 64)]
        oimgs, _, _ = g_forward(sd, z[:16], emb[:16], eps[:16], cfg, False)
    assert [tuple(t.shape) for t in full] == [(B, 3, 64, 64), (B, 3, 128, 128), (B, 3, 256, 256)]
    for lvl in range(3):
        cat = torch.cat([p[lvl] for p in parts], 0)
        assert rel(full[lvl], cat) < 1.5e-2, (lvl, rel(full[lvl], cat))
        assert rel(full[lvl][:16], oimgs[lvl]) < FWD_TOL, (lvl, rel(full[lvl][:16], oimgs[lvl]))


@pytest.mark.parametrize("which,B", [(0, 8), (1, 6), (2, 4)])
def test_d_forward_backward(which, B):
    cfg = Cfg()
    fp32_strict()
    net, sd = make_d(cfg, which, seed=2)
    for k in sd:
        if is_param(k):
            sd[k].requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    S = 64 * 2 ** which
    base = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
    c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    # non-leaf image and c (the G-step situation): gradients must flow to both
    img = base.clone().requires_grad_(True)
    c = c0.clone().requires_grad_(True)
    oimg = base.clone().requires_grad_(True)
    oc = c0.clone().requires_grad_(True)
    (cond, uncond), x_imm = net(img * 1.0, c * 1.0)
    (ocond, ouncond), ox = d_forward(sd, oimg * 1.0, oc * 1.0, which, cfg, True)
    pairs = [("cond", rel(cond, ocond)), ("uncond", rel(uncond, ouncond)), ("x_immediate", rel(x_imm, ox))]
    ok, msg = report(pairs, D_FWD_TOL)
    assert ok, msg
    r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
    r3 = torch.randn(ox.shape, generator=g).cuda() * 0.01
    ((cond * r1).sum() + (uncond * r2).sum() + (x_imm * r3).sum()).backward()
    ((ocond * r1).sum() + (ouncond * r2).sum() + (ox * r3).sum()).backward()
    coss = [(k, cos(p.grad, sd[k].grad)) for k, p in net.named_parameters()]
    coss += [("d_img", cos(img.grad, oimg.grad)), ("d_c", cos(c.grad, oc.grad))]
    bad = [(k, v) for k, v in coss if not v > D_GRAD_COS]
    assert not bad, bad
    norms = [(k, float(p.grad.norm() / (sd[k].grad.norm() + 1e-30))) for k, p in net.named_parameters()]
    assert all(0.9 < v < 1.1 for _, v in norms), [kv for kv in norms if not 0.9 < kv[1] < 1.1]
    st = net.state_dict()
    ok, msg = report([(k, rel(st[k].float(), sd[k].float())) for k in sd if "running" in k], D_FWD_TOL)
    assert ok, msg


def test_cpu_tensors_are_rejected():
    cfg = Cfg(BRANCH_NUM=1)
    net, _ = make_g(cfg)
    with pytest.raises(RuntimeError):
        net(torch.randn(2, cfg.Z_DIM), torch.randn(2, cfg.TEXT_DIM))


@pytest.mark.parametrize("which,B", [(0, 8), (2, 4)])
def test_d_batched_groups_equal_separate_passes(which, B):
    """train_Dnet's real / wrong / fake passes (trainer.py:390-392) run as ONE pass over the concatenated 3B samples
    with per-sub-batch BatchNorm statistics must reproduce the three separate passes: logits, every parameter
    gradient, running statistics (three momentum updates, in order) and num_batches_tracked (+3). Both arms run the
    same bf16 kernels; the split-K factor of the deep layers depends on the GEMM row count, so fp32 summation order
    differs, a few bf16 roundings flip and the small-batch BatchNorm of the 4x4 maps amplifies them (same effect as the
    run-to-run tolerance in test_gpu_train_step.py): logits rel <= 8e-3 (measured <= 4.2e-3), gradient cosine >= 0.99
    (measured 0.9953-0.99999; the LeakyReLU mask flips behind it are the ones described in the module docstring)."""
    from sg2b200.nets import GradSink
    cfg = Cfg()
    fp32_strict()
    netA, _ = make_d(cfg, which, seed=4)
    netB, _ = make_d(cfg, which, seed=4)
    g = torch.Generator().manual_seed(11)
    S = 64 * 2 ** which
    imgs = [(torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda() for _ in range(3)]
    c = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    dco = [torch.randn(B, generator=g).cuda() * 0.1 for _ in range(3)]
    dun = [torch.randn(B, generator=g).cuda() * 0.1 for _ in range(3)]
    EA, EB = netA.engine(), netB.engine()
    sinkA, tapes, pa = GradSink(), [], []
    for k in range(3):
        cond, uncond, _, T = EA.forward(imgs[k], c, True)
        tapes.append(T)
        pa.append((cond, uncond))
    for k in range(3):
        EA.backward(tapes[k], dco[k], dun[k], None, False, False, True, sinkA)
    gA = sinkA.finish()
    sinkB = GradSink()
    cond3, uncond3, _, T3 = EB.forward(torch.cat(imgs, 0), c.repeat(3, 1), True, groups=3)
    EB.backward(T3, torch.cat(dco), torch.cat(dun), None, False, False, True, sinkB)
    gB = sinkB.finish()
    pairs = [(f"cond{k}", rel(cond3[k * B:(k + 1) * B], pa[k][0])) for k in range(3)]
    pairs += [(f"uncond{k}", rel(uncond3[k * B:(k + 1) * B], pa[k][1])) for k in range(3)]
    ok, msg = report(pairs, 8e-3)
    assert ok, msg
    namesA = dict(netA.named_parameters())
    namesB = dict(netB.named_parameters())
    coss = [(k, cos(gB[namesB[k]], gA[namesA[k]])) for k in namesA]
    assert all(v > 0.99 for _, v in coss), [kv for kv in coss if not kv[1] > 0.99]
    stA, stB = netA.state_dict(), netB.state_dict()
    ok, msg = report([(k, rel(stB[k].float(), stA[k].float())) for k in stA if "running" in k], 1e-3)
    assert ok, msg
    assert all(int(stA[k]) == int(stB[k]) == 3 for k in stA if "num_batches" in k)
