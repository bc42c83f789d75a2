"""Measured parity table of the CUDA path (not pytest): every number the tolerances in tests/ are derived from.

    python tools/parity_report.py [--out gpurun_out/parity_r02.json] [--quick]

Sections: (A) run-to-run reproducibility of the fused step, (B) CUDA-graph CapturedStep vs eager FusedTrainer.step,
(C) fused step vs the fp32 oracle and vs the bf16-emulating oracle (losses, every gradient, updated weights, BN
buffers), (D) G / D forward per tensor. Uses the oracle as the checker only.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.stackgan_oracle import Cfg, d_forward, emulate_bf16, g_forward, is_param, param_keys  # noqa: E402
from tests.parity_util import (bucket_grads, build_trainer_and_oracles, cos, fp32_strict, loss_vector, make_d, make_g,  # noqa: E402
                               oracle_grads, oracle_step, rel, snapshot_diff, train_batch)


def summarise(pairs):
    """pairs: list of (name, rel, cos)."""
    rels = sorted(p[1] for p in pairs)
    worst = max(pairs, key=lambda p: p[1])
    lowc = min(pairs, key=lambda p: p[2])
    return {"n": len(pairs), "rel_max": worst[1], "rel_max_at": worst[0], "rel_median": rels[len(rels) // 2],
            "cos_min": lowc[2], "cos_min_at": lowc[0]}


def section_repro(branches, B, steps=2):
    from sg2b200 import trainer
    cfg, ocfg, netG, netsD, tr, _ = build_trainer_and_oracles(branches, seed=0, n_oracles=0)
    batches = [train_batch(cfg, B, 300 + s) for s in range(steps)]
    snap0 = tr.snapshot()

    def eager():
        tr.restore(snap0)
        ls = [tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).clone() for b in batches]
        torch.cuda.synchronize()
        return tr.snapshot(), torch.stack(ls)

    sA, lA = eager()
    sB, lB = eager()
    out = {"eager_vs_eager": dict(zip(("max_rel", "bitwise"), snapshot_diff(sA, sB))) | {"losses_bitwise": bool(torch.equal(lA, lB))}}
    for conc in (True, False):
        tr.concurrent = conc
        cap = trainer.CapturedStep(tr, B, warmup=2, draw_noise=False)
        b = batches[0]
        cap.load(b["emb"], b["real"], b["wrong"], b["labels"], z=b["z"], eps=b["eps"])
        cap.capture()
        res = []
        for rep in range(2):
            tr.restore(snap0)
            ls = []
            for b in batches:
                cap.load(b["emb"], b["real"], b["wrong"], b["labels"], z=b["z"], eps=b["eps"])
                ls.append(cap.replay().clone())
            torch.cuda.synchronize()
            res.append((tr.snapshot(), torch.stack(ls)))
        d, same = snapshot_diff(res[0][0], sA)
        out[f"captured_concurrent={conc}_vs_eager"] = {"max_rel": d, "bitwise": same,
                                                       "losses_max_rel": rel(res[0][1], lA),
                                                       "losses_bitwise": bool(torch.equal(res[0][1], lA))}
        d2, same2 = snapshot_diff(res[0][0], res[1][0])
        out[f"captured_concurrent={conc}_replay_vs_replay"] = {"max_rel": d2, "bitwise": same2}
        del cap
    tr.concurrent = True
    return out


def _flat(d, prefix):
    return torch.cat([v.detach().double().flatten() for k, v in d.items() if k.startswith(prefix)])


def section_step(branches, B, precision="bf16", lr=0.0):
    """One fused step against the FLOAT64 oracle (truth); the fp32 oracle (PyTorch CUDA fp32, the reference's own
    arithmetic) and the bf16-emulating oracle (ideal bf16 storage) are measured against the same truth as yardsticks.
    lr = 0: a comparison of gradients (the G step sees identical D weights in every arm)."""
    from oracle.stackgan_oracle import OracleTrainer
    from sg2b200 import config
    from tests.parity_util import f64_state
    config.set_precision(precision)
    try:
        cfg, ocfg, netG, netsD, tr, (o32, oq) = build_trainer_and_oracles(branches, seed=0, n_oracles=2, lr=lr)
        o64 = OracleTrainer(ocfg, f64_state({k: v.detach() for k, v in o32.g.items()}),
                            [f64_state({k: v.detach() for k, v in d.items()}) for d in o32.ds], device="cuda")
        b = train_batch(cfg, B, 11)
        losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu().tolist()
        rt = oracle_step(o64, b, keep_grads=True)
        r32 = oracle_step(o32, b, keep_grads=True)
        with emulate_bf16():
            rq = oracle_step(oq, b, keep_grads=True)
        names = [f"errD{i}" for i in range(branches)] + ["errG_total", "kl", "cal"]
        lt = loss_vector(rt)
        relv = lambda xs: [abs(a - t) / (abs(t) + 1e-12) for a, t in zip(xs, lt)]
        rec = {"precision": precision, "lr": lr,
               "loss_rel_vs_f64": {"names": names, "ours": relv(losses), "fp32_oracle": relv(loss_vector(r32)),
                                   "bf16_emulating_oracle": relv(loss_vector(rq))}}
        truth = oracle_grads(rt)
        arms = {"ours": bucket_grads(tr), "fp32_oracle": oracle_grads(r32), "bf16_emulating_oracle": oracle_grads(rq)}
        nets_ = ["G"] + [f"D{i}" for i in range(branches)]
        rec["grad_flat_rel_vs_f64"] = {arm: {n: rel(_flat(g, n + "."), _flat(truth, n + ".")) for n in nets_} for arm, g in arms.items()}
        rec["grad_flat_cos_vs_f64"] = {arm: {n: cos(_flat(g, n + "."), _flat(truth, n + ".")) for n in nets_} for arm, g in arms.items()}
        rec["grad_per_tensor_vs_f64"] = {arm: summarise([(k, rel(g[k], truth[k]), cos(g[k], truth[k])) for k in truth]) for arm, g in arms.items()}
        sd = netG.state_dict()
        rec["G_running_stats_rel_max_vs_f64"] = max(rel(sd[k].float(), o64.g[k].float()) for k in o64.g if "running" in k)
        return rec
    finally:
        config.set_precision("bf16")


def section_forward():
    out = {}
    fp32_strict()
    for branches, B in ((1, 8), (3, 4)):
        cfg = Cfg(BRANCH_NUM=branches)
        net, sd = make_g(cfg, seed=1)
        g = torch.Generator().manual_seed(3)
        z = torch.randn(B, cfg.Z_DIM, generator=g).cuda()
        emb = torch.randn(B, cfg.TEXT_DIM, generator=g).cuda()
        eps = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
        from tests.parity_util import f64_state
        sdq = {k: v.clone() for k, v in sd.items()}
        sdt = f64_state(sd)
        for s in (sd, sdq, sdt):
            for k in s:
                if is_param(k):
                    s[k].requires_grad_(True)
        imgs, mu, logvar = net(z, emb, eps=eps)
        o = g_forward(sd, z, emb, eps, cfg, True)
        t = g_forward(sdt, z.double(), emb.double(), eps.double(), cfg, True)
        with emulate_bf16():
            q = g_forward(sdq, z, emb, eps, cfg, True)
        rs = [torch.randn(i.shape, generator=g).cuda() for i in o[0]]
        rmu, rlv = torch.randn(mu.shape, generator=g).cuda(), torch.randn(mu.shape, generator=g).cuda()
        L = lambda t_: sum((a * r).sum() for a, r in zip(t_[0], rs)) + (t_[1] * rmu).sum() + (t_[2] * rlv).sum()
        L((imgs, mu, logvar)).backward()
        L(o).backward()
        L(t).backward()
        with emulate_bf16():
            L(q).backward()
        rec = {"img_rel_vs_f64": {"ours": [rel(a, b) for a, b in zip(imgs, t[0])], "fp32_oracle": [rel(a, b) for a, b in zip(o[0], t[0])],
                                  "bf16_emulating_oracle": [rel(a, b) for a, b in zip(q[0], t[0])]}, "mu_rel": rel(mu, t[1])}
        for tag, s_ in (("ours", None), ("fp32_oracle", sd), ("bf16_emulating_oracle", sdq)):
            gr = {k: (p.grad if s_ is None else s_[k].grad) for k, p in net.named_parameters()}
            rec[f"grads_vs_f64_{tag}"] = summarise([(k, rel(gr[k], sdt[k].grad), cos(gr[k], sdt[k].grad)) for k in gr])
            rec[f"grads_vs_f64_{tag}"]["flat_rel"] = rel(torch.cat([gr[k].double().flatten() for k in gr]), torch.cat([sdt[k].grad.flatten() for k in gr]))
        out[f"G_branches{branches}_B{B}"] = rec
    for which, B in ((0, 8), (1, 6), (2, 4)):
        cfg = Cfg()
        net, sd = make_d(cfg, which, seed=2)
        from tests.parity_util import f64_state
        sdq = {k: v.clone() for k, v in sd.items()}
        sdt = f64_state(sd)
        for s in (sd, sdq, sdt):
            for k in s:
                if is_param(k):
                    s[k].requires_grad_(True)
        g = torch.Generator().manual_seed(5)
        S = 64 * 2 ** which
        base = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
        c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
        leaves = [(base.clone().requires_grad_(True), c0.clone().requires_grad_(True)) for _ in range(3)]
        lt = (base.double().requires_grad_(True), c0.double().requires_grad_(True))
        (cond, uncond), x_imm = net(leaves[0][0] * 1.0, leaves[0][1] * 1.0)
        (oc, ou), ox = d_forward(sd, leaves[1][0] * 1.0, leaves[1][1] * 1.0, which, cfg, True)
        (tc, tu), tx = d_forward(sdt, lt[0] * 1.0, lt[1] * 1.0, which, cfg, True)
        with emulate_bf16():
            (qc, qu), qx = d_forward(sdq, leaves[2][0] * 1.0, leaves[2][1] * 1.0, which, cfg, True)
        r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
        r3 = torch.randn(ox.shape, generator=g).cuda() * 0.01
        ((cond * r1).sum() + (uncond * r2).sum() + (x_imm * r3).sum()).backward()
        ((oc * r1).sum() + (ou * r2).sum() + (ox * r3).sum()).backward()
        ((tc * r1).sum() + (tu * r2).sum() + (tx * r3).sum()).backward()
        with emulate_bf16():
            ((qc * r1).sum() + (qu * r2).sum() + (qx * r3).sum()).backward()
        rec = {"fwd_rel_vs_f64": {"names": ["cond", "uncond", "x_immediate"], "ours": [rel(cond, tc), rel(uncond, tu), rel(x_imm, tx)],
                                  "fp32_oracle": [rel(oc, tc), rel(ou, tu), rel(ox, tx)],
                                  "bf16_emulating_oracle": [rel(qc, tc), rel(qu, tu), rel(qx, tx)]}}
        tg = {k: sdt[k].grad for k, _ in net.named_parameters()} | {"d_img": lt[0].grad, "d_c": lt[1].grad}
        for tag, s, lv in (("ours", None, leaves[0]), ("fp32_oracle", sd, leaves[1]), ("bf16_emulating_oracle", sdq, leaves[2])):
            gr = {k: (p.grad if s is None else s[k].grad) for k, p in net.named_parameters()} | {"d_img": lv[0].grad, "d_c": lv[1].grad}
            rec[f"grads_vs_f64_{tag}"] = summarise([(k, rel(gr[k], tg[k]), cos(gr[k], tg[k])) for k in gr])
            rec[f"grads_vs_f64_{tag}"]["flat_rel"] = rel(torch.cat([gr[k].double().flatten() for k in gr]), torch.cat([tg[k].flatten() for k in gr]))
        out[f"D{which}_B{B}"] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    res = {}
    res["forward"] = section_forward()
    res["repro_3stage_B6"] = section_repro(3, 6)
    for prec in ("bf16", "fp32"):
        res[f"step_1stage_B8_{prec}"] = section_step(1, 8, prec)
        if not args.quick:
            res[f"step_3stage_B24_{prec}"] = section_step(3, 24, prec)
    if not args.quick:
        res["step_3stage_B24_bf16_lr2e-4"] = section_step(3, 24, "bf16", lr=2e-4)
        res["repro_3stage_B24"] = section_repro(3, 24)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
