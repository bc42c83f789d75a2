"""GPU parity of the bf16 path: sg2b200 G_NET / D_NET* (CUDA kernels through the C ABI) against the oracle on identical
weights and inputs.

Metric: per-tensor relative error ||a - b|| / ||b|| against a FLOAT64 run of the oracle (the ground truth of the
reference's arithmetic). The CUDA path stores every activation and every backward tensor in bf16 (fp32 accumulation).
What that storage format costs is measured, not assumed: the oracle is run a second time with bf16 rounding at exactly
the points where the CUDA path stores a tensor, forward and backward (`emulate_bf16`), i.e. an IDEAL implementation of
bf16 storage in otherwise exact fp32 arithmetic. The test then is: the kernels are as close to the truth as that ideal
implementation (factor YARD, plus a small floor for tensors whose error is itself tiny) — anything the kernels did
wrong on top of the storage format would show up as a ratio > 1. Measured ratios on B200: 0.95 - 1.05 for every image,
logit vector and flat gradient (profiles/r02_parity.md).

Absolute numbers for the record (B200, profiles/r02_parity.md): images 64 / 128 / 256 px 0.8e-2 / 1.2e-2 / 1.8e-2 after
8 / 14 / 20 bf16 conv + BN layers; G gradients 1-2e-2 flat; D logits 1-6e-3, x_immediate <= 1.2e-2; D gradients 6-11e-2
flat (cosine >= 0.986): gradients through LeakyReLU are discontinuous at 0, so their relative error scales like the
SQUARE ROOT of the storage precision (tests/test_gpu_precise.py explains and shows the same for fp32: 1e-3, not 1e-7).
The north star's 1e-2 holds per layer (tests/test_gpu_kernels.py: <= 5e-3), for the first stage's image and for every
loss (<= 1e-3); the fp32-accurate mode (tests/test_gpu_precise.py) is the one that reaches 1e-4."""
import pytest
import torch

from oracle.stackgan_oracle import Cfg, d_forward, emulate_bf16, g_forward, is_param
from tests.parity_util import f64_state, fp32_strict, make_d, make_g, rel, report

pytestmark = pytest.mark.gpu
YARD = 1.3             # ours-vs-truth <= YARD x (ideal bf16 storage)-vs-truth (+ floor); measured 0.95 - 1.05
FWD_TOL = 2.5e-2       # absolute backstop: G images after up to 20 bf16 layers (measured <= 1.84e-2)
D_FWD_TOL = 2e-2       # absolute backstop: D logits / x_immediate (measured <= 1.23e-2)
G_GRAD_TOL = 7e-2      # absolute backstop, per tensor (measured <= 5.7e-2, flat 2e-2, cosine >= 0.998)
D_GRAD_COS = 0.98      # absolute backstop (measured cosine >= 0.986)


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def _flat(ts):
    return torch.cat([t.detach().double().flatten() for t in ts])


def _track(sd):
    for k in sd:
        if is_param(k):
            sd[k].requires_grad_(True)
    return sd


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
    sdq = _track({k: v.clone() for k, v in sd.items()})          # ideal bf16 storage
    sdt = _track(f64_state(sd))                                   # truth
    imgs, mu, logvar = net(z, emb, eps=eps)
    timgs, tmu, tlogvar = g_forward(sdt, z.double(), emb.double(), eps.double(), cfg, True)
    with emulate_bf16():
        qimgs, qmu, qlogvar = g_forward(sdq, z, emb, eps, cfg, True)
    for i, (a, q, t) in enumerate(zip(imgs, qimgs, timgs)):
        e, eq = rel(a, t), rel(q, t)
        assert e <= YARD * eq + 1e-3 and e <= FWD_TOL, (f"img{i}", e, eq)
    assert rel(mu, tmu) < 1e-5 and rel(logvar, tlogvar) < 1e-5     # CA_NET runs in fp32
    # BN running statistics updated like nn.BatchNorm (momentum 0.1, unbiased var)
    st = net.state_dict()
    ok, msg = report([(k, rel(st[k].float(), sdt[k].float())) for k in sdt if "running" in k], FWD_TOL)
    assert ok, msg
    assert all(int(st[k]) == int(sdt[k]) for k in sdt if "num_batches" in k)
    # backward: random cotangents on every output
    rs = [torch.randn(i.shape, generator=g).cuda() for i in timgs]
    rmu, rlv = torch.randn(mu.shape, generator=g).cuda(), torch.randn(mu.shape, generator=g).cuda()
    L = lambda t3: sum((a * r).sum() for a, r in zip(t3[0], rs)) + (t3[1] * rmu).sum() + (t3[2] * rlv).sum()
    L((imgs, mu, logvar)).backward()
    L((timgs, tmu, tlogvar)).backward()
    with emulate_bf16():
        L((qimgs, qmu, qlogvar)).backward()
    names = [k for k, _ in net.named_parameters()]
    ours = {k: p.grad for k, p in net.named_parameters()}
    e, eq = rel(_flat(ours[k] for k in names), _flat(sdt[k].grad for k in names)), rel(_flat(sdq[k].grad for k in names), _flat(sdt[k].grad for k in names))
    assert e <= YARD * eq + 1e-3, (e, eq)
    ok, msg = report([(k, rel(ours[k], sdt[k].grad)) for k in names], G_GRAD_TOL)
    assert ok, msg
    assert min(cos(ours[k], sdt[k].grad) for k in names) > 0.998


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
    sdq = _track({k: v.clone() for k, v in sd.items()})
    sdt = _track(f64_state(sd))
    g = torch.Generator().manual_seed(5)
    S = 64 * 2 ** which
    base = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
    c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    # non-leaf image and c (the G-step situation): gradients must flow to both
    img, c = base.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    qimg, qc = base.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    timg, tc = base.double().requires_grad_(True), c0.double().requires_grad_(True)
    (cond, uncond), x_imm = net(img * 1.0, c * 1.0)
    (tcond, tuncond), tx = d_forward(sdt, timg * 1.0, tc * 1.0, which, cfg, True)
    with emulate_bf16():
        (qcond, quncond), qx = d_forward(sdq, qimg * 1.0, qc * 1.0, which, cfg, True)
    for name, a, q, t in (("cond", cond, qcond, tcond), ("uncond", uncond, quncond, tuncond), ("x_immediate", x_imm, qx, tx)):
        e, eq = rel(a, t), rel(q, t)
        # B-element logit vectors: the two bf16 realisations differ by up to 2x on a handful of numbers
        assert e <= 2.0 * eq + 2e-3 and e <= D_FWD_TOL, (name, e, eq)
    r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
    r3 = torch.randn(tx.shape, generator=g).cuda() * 0.01
    ((cond * r1).sum() + (uncond * r2).sum() + (x_imm * r3).sum()).backward()
    ((tcond * r1).sum() + (tuncond * r2).sum() + (tx * r3).sum()).backward()
    with emulate_bf16():
        ((qcond * r1).sum() + (quncond * r2).sum() + (qx * r3).sum()).backward()
    names = [k for k, _ in net.named_parameters()]
    ours = {k: p.grad for k, p in net.named_parameters()} | {"d_img": img.grad, "d_c": c.grad}
    emu = {k: sdq[k].grad for k in names} | {"d_img": qimg.grad, "d_c": qc.grad}
    truth = {k: sdt[k].grad for k in names} | {"d_img": timg.grad, "d_c": tc.grad}
    e, eq = rel(_flat(ours[k] for k in truth), _flat(truth.values())), rel(_flat(emu[k] for k in truth), _flat(truth.values()))
    assert e <= YARD * eq + 1e-3, (e, eq)
    bad = [(k, cos(ours[k], truth[k])) for k in truth if not cos(ours[k], truth[k]) > D_GRAD_COS]
    assert not bad, bad
    norms = [(k, float(ours[k].norm() / (truth[k].norm() + 1e-30))) for k in names]
    assert all(0.9 < v < 1.1 for _, v in norms), [kv for kv in norms if not 0.9 < kv[1] < 1.1]
    st = net.state_dict()
    ok, msg = report([(k, rel(st[k].float(), sdt[k].float())) for k in sdt if "running" in k], D_FWD_TOL)
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
