"""GPU parity of the path bench.py times: the whole train step replayed from ONE CUDA graph (trainer.CapturedStep:
multi-stream capture, per-stream scratch arenas rewound during capture, per-layer Adam + operand re-packs on the side
streams, device-side Adam step counter) against the eager FusedTrainer.step and against the oracle, over several
steps with changing inputs and given noise. Reference loop: trainer.py:529-572, train_Dnet 375-427, train_Gnet 429-489.

The reductions of the step are reproducible (ops.DETERMINISTIC: ordered slab sums for split-K / wgrad, fp64 atomics over
fixed-order per-block partials elsewhere), so eager and replayed steps that launch the same kernels must agree to the
last bit up to the 2^-53 order dependence of the fp64 atomics: asserted as relative difference <= 1e-6 on every
parameter, Adam moment, EMA weight and BatchNorm buffer (measured: bit-identical, profiles/r02_parity.md)."""
import pytest
import torch

from oracle.stackgan_oracle import emulate_bf16
from tests.parity_util import (build_trainer_and_oracles, loss_vector, oracle_step, rel, snapshot_diff, train_batch)

pytestmark = pytest.mark.gpu
STEPS = 3


def _eager(tr, snap0, batches):
    tr.restore(snap0)
    ls = [tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).clone() for b in batches]
    torch.cuda.synchronize()
    return tr.snapshot(), torch.stack(ls)


def _replayed(tr, cap, snap0, batches):
    tr.restore(snap0)
    ls = []
    for b in batches:
        cap.load(b["emb"], b["real"], b["wrong"], b["labels"], z=b["z"], eps=b["eps"])
        ls.append(cap.replay().clone())
    torch.cuda.synchronize()
    return tr.snapshot(), torch.stack(ls)


@pytest.mark.parametrize("branches,B,concurrent", [(3, 24, True), (3, 24, False), (1, 8, True)])   # (3, 24) = configs[1]
def test_captured_step_replays_equal_eager_steps(branches, B, concurrent):
    from sg2b200 import ops, trainer
    cfg, ocfg, netG, netsD, tr, (orc,) = build_trainer_and_oracles(branches, seed=0, n_oracles=1)
    tr.concurrent = concurrent
    batches = [train_batch(cfg, B, 500 + s) for s in range(STEPS)]
    snap0 = tr.snapshot()
    s_eager, l_eager = _eager(tr, snap0, batches)
    cap = trainer.CapturedStep(tr, B, warmup=2, draw_noise=False)
    b0 = batches[0]
    cap.load(b0["emb"], b0["real"], b0["wrong"], b0["labels"], z=b0["z"], eps=b0["eps"])
    cap.capture()                                       # runs warm-up steps: the state is restored before replaying
    s_cap, l_cap = _replayed(tr, cap, snap0, batches)
    tol = 1e-6 if ops.DETERMINISTIC else 5e-2
    d, bitwise = snapshot_diff(s_cap, s_eager)
    assert d <= tol, (d, bitwise)
    assert rel(l_cap, l_eager) <= tol, (l_cap.tolist(), l_eager.tolist())
    # a second round of replays from the same state reproduces the first (no stale slot / pack / counter in the graph)
    s_cap2, l_cap2 = _replayed(tr, cap, snap0, batches)
    d2, _ = snapshot_diff(s_cap2, s_cap)
    assert d2 <= tol, d2
    # Adam's device-side step counter advanced once per replayed step in every bucket
    for bkt in [tr.bG] + tr.bD:
        assert int(bkt.step) == int(snap0["buckets"][0]["step"]) + STEPS
    # and the replayed losses follow the bf16-emulating oracle over the same steps (tolerance: test_gpu_train_step.py)
    with emulate_bf16():
        ref = torch.tensor([loss_vector(oracle_step(orc, b)) for b in batches])
    got = l_cap.cpu()
    assert ((got - ref).abs() <= 3e-2 * ref.abs() + 2e-3).all(), (got.tolist(), ref.tolist())


def test_fused_step_is_reproducible_run_to_run():
    from sg2b200 import ops
    if not ops.DETERMINISTIC:
        pytest.skip("SG2_DETERMINISTIC=0")
    cfg, ocfg, netG, netsD, tr, _ = build_trainer_and_oracles(3, seed=1, n_oracles=0)
    batches = [train_batch(cfg, 6, 700 + s) for s in range(2)]
    snap0 = tr.snapshot()
    sA, lA = _eager(tr, snap0, batches)
    sB, lB = _eager(tr, snap0, batches)
    d, bitwise = snapshot_diff(sA, sB)
    assert d <= 1e-6 and rel(lA, lB) <= 1e-6, (d, bitwise)


def test_ema_swap_through_data_writes_is_seen_by_the_kernels():
    """the reference's load_params writes `p.data.copy_` (trainer.py:78-80), which bumps no version counter.
    Module-API flow of trainer.py:592-601: training forward, EMA weights swapped in through `.data`, snapshot forward
    under no_grad, weights swapped back, next training forward — every forward must see the weights that are in the
    parameters at that time."""
    from oracle.stackgan_oracle import Cfg
    from tests.parity_util import make_g
    cfg = Cfg(BRANCH_NUM=2)
    net, _ = make_g(cfg, seed=4)
    g = torch.Generator().manual_seed(1)
    z, emb = torch.randn(4, cfg.Z_DIM, generator=g).cuda(), torch.randn(4, cfg.TEXT_DIM, generator=g).cuda()
    eps = torch.randn(4, cfg.EMBEDDING_DIM, generator=g).cuda()
    net.eval()                                             # same BN statistics for every call below
    live = [t.detach().clone() for t in net(z, emb, eps=eps)[0]]           # grad mode: packs are cached on versions
    backup = [p.data.clone() for p in net.parameters()]
    avg = [p.data + 0.05 * torch.randn_like(p.data) for p in net.parameters()]
    for p, a in zip(net.parameters(), avg):                # == the reference's load_params(netG, avg_param_G)
        p.data.copy_(a)
    with torch.no_grad():
        ema = [t.clone() for t in net(z, emb, eps=eps)[0]]
    for p, a in zip(net.parameters(), backup):             # == load_params(netG, backup_para)
        p.data.copy_(a)
    back = [t.detach().clone() for t in net(z, emb, eps=eps)[0]]
    assert all(rel(a, b) > 1e-3 for a, b in zip(ema, live))
    assert all(torch.equal(a, b) for a, b in zip(back, live))


def test_fused_trainer_ema_swap_and_external_writes():
    """FlatBucket.swap_ema() (the fused trainer's form of the snapshot swap) and utils.load_params / FlatBucket.refresh()
    after writes from outside the fused kernels."""
    from sg2b200 import utils
    cfg, ocfg, netG, netsD, tr, _ = build_trainer_and_oracles(1, seed=2, n_oracles=0)
    b = train_batch(cfg, 8, 900)
    for s in range(2):
        tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"])
    netG.eval()
    with torch.no_grad():
        live = netG(b["z"], b["emb"], eps=b["eps"])[0][0].clone()
        tr.bG.swap_ema()
        ema = netG(b["z"], b["emb"], eps=b["eps"])[0][0].clone()
        tr.bG.swap_ema()
        back = netG(b["z"], b["emb"], eps=b["eps"])[0][0].clone()
        assert rel(ema, live) > 1e-5 and torch.equal(back, live)
        utils.load_params(netG, tr.bG.ema_params())        # explicit API on bucket-backed parameters
        ema2 = netG(b["z"], b["emb"], eps=b["eps"])[0][0].clone()
        assert torch.equal(ema2, ema)
    netG.train()
