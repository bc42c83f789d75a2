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


def section_step(branches, B, steps=1):
    cfg, ocfg, netG, netsD, tr, (o32, oq) = build_trainer_and_oracles(branches, seed=0, n_oracles=2)
    out = {"steps": []}
    for s in range(steps):
        b = train_batch(cfg, B, 11 + s)
        w0 = {k: v.detach().clone() for k, v in netG.state_dict().items() if is_param(k)}
        losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu().tolist()
        r32 = oracle_step(o32, b, keep_grads=True)
        with emulate_bf16():
            rq = oracle_step(oq, b, keep_grads=True)
        names = [f"errD{i}" for i in range(branches)] + ["errG_total", "kl", "cal"]
        rec = {"losses": {n: {"ours": a, "fp32": x, "emu": y, "rel_fp32": abs(a - x) / (abs(x) + 1e-12),
                              "rel_emu": abs(a - y) / (abs(y) + 1e-12)}
                          for n, a, x, y in zip(names, losses, loss_vector(r32), loss_vector(rq))}}
        ours = bucket_grads(tr)
        for tag, ref in (("fp32", oracle_grads(r32)), ("emu", oracle_grads(rq))):
            per_net = {}
            for k, g in ours.items():
                net = k.split(".")[0]
                per_net.setdefault(net, []).append((k, rel(g, ref[k]), cos(g, ref[k])))
            rec[f"grads_vs_{tag}"] = {net: summarise(v) for net, v in per_net.items()}
            # whole-network gradient (all parameters concatenated)
            for net in per_net:
                a = torch.cat([ours[k].flatten() for k in ours if k.startswith(net + ".")])
                r = torch.cat([ref[k].flatten() for k in ours if k.startswith(net + ".")])
                rec[f"grads_vs_{tag}"][net]["flat_rel"] = rel(a, r)
                rec[f"grads_vs_{tag}"][net]["flat_cos"] = cos(a, r)
        # emu oracle vs fp32 oracle: the quantisation gap itself
        per = [(k, rel(oracle_grads(rq)[k], oracle_grads(r32)[k]), cos(oracle_grads(rq)[k], oracle_grads(r32)[k])) for k in ours]
        rec["emu_vs_fp32_grads"] = summarise(per)
        # updated weights: how far apart are the Adam steps (units of lr)
        lr = ocfg.LR_G
        sd = netG.state_dict()
        upd = []
        for k in param_keys(o32.g):
            d_ours, d_ref = sd[k].detach() - w0[k], o32.g[k].detach() - w0[k]
            upd.append((k, float((d_ours - d_ref).abs().mean()) / lr, cos(d_ours, d_ref)))
        rec["G_update_vs_fp32"] = {"mean_abs_diff_over_lr_max": max(u[1] for u in upd), "cos_min": min(u[2] for u in upd)}
        bn = [(k, rel(sd[k].float(), o32.g[k].float())) for k in o32.g if "running" in k]
        rec["G_running_stats_rel_max_fp32"] = max(v for _, v in bn)
        out["steps"].append(rec)
    return out


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
        sdq = {k: v.clone() for k, v in sd.items()}
        sds = [sd, sdq]
        for s in sds:
            for k in s:
                if is_param(k):
                    s[k].requires_grad_(True)
        imgs, mu, logvar = net(z, emb, eps=eps)
        o = g_forward(sd, z, emb, eps, cfg, True)
        with emulate_bf16():
            q = g_forward(sdq, z, emb, eps, cfg, True)
        rs = [torch.randn(i.shape, generator=g).cuda() for i in o[0]]
        rmu, rlv = torch.randn(mu.shape, generator=g).cuda(), torch.randn(mu.shape, generator=g).cuda()
        L = lambda t: sum((a * r).sum() for a, r in zip(t[0], rs)) + (t[1] * rmu).sum() + (t[2] * rlv).sum()
        L((imgs, mu, logvar)).backward()
        L(o).backward()
        with emulate_bf16():
            L(q).backward()
        rec = {"img_rel_fp32": [rel(a, b) for a, b in zip(imgs, o[0])], "img_rel_emu": [rel(a, b) for a, b in zip(imgs, q[0])],
               "emu_vs_fp32_img": [rel(a, b) for a, b in zip(q[0], o[0])], "mu_rel": rel(mu, o[1])}
        rec["grads_vs_fp32"] = summarise([(k, rel(p.grad, sd[k].grad), cos(p.grad, sd[k].grad)) for k, p in net.named_parameters()])
        rec["grads_vs_emu"] = summarise([(k, rel(p.grad, sdq[k].grad), cos(p.grad, sdq[k].grad)) for k, p in net.named_parameters()])
        out[f"G_branches{branches}_B{B}"] = rec
    for which, B in ((0, 8), (1, 6), (2, 4)):
        cfg = Cfg()
        net, sd = make_d(cfg, which, seed=2)
        sdq = {k: v.clone() for k, v in sd.items()}
        for s in (sd, sdq):
            for k in s:
                if is_param(k):
                    s[k].requires_grad_(True)
        g = torch.Generator().manual_seed(5)
        S = 64 * 2 ** which
        base = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
        c0 = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
        leaves = [(base.clone().requires_grad_(True), c0.clone().requires_grad_(True)) for _ in range(3)]
        (cond, uncond), x_imm = net(leaves[0][0] * 1.0, leaves[0][1] * 1.0)
        (oc, ou), ox = d_forward(sd, leaves[1][0] * 1.0, leaves[1][1] * 1.0, which, cfg, True)
        with emulate_bf16():
            (qc, qu), qx = d_forward(sdq, leaves[2][0] * 1.0, leaves[2][1] * 1.0, which, cfg, True)
        r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
        r3 = torch.randn(ox.shape, generator=g).cuda() * 0.01
        ((cond * r1).sum() + (uncond * r2).sum() + (x_imm * r3).sum()).backward()
        ((oc * r1).sum() + (ou * r2).sum() + (ox * r3).sum()).backward()
        with emulate_bf16():
            ((qc * r1).sum() + (qu * r2).sum() + (qx * r3).sum()).backward()
        rec = {"fwd_rel_fp32": [rel(cond, oc), rel(uncond, ou), rel(x_imm, ox)],
               "fwd_rel_emu": [rel(cond, qc), rel(uncond, qu), rel(x_imm, qx)]}
        for tag, s, lv in (("fp32", sd, leaves[1]), ("emu", sdq, leaves[2])):
            pairs = [(k, rel(p.grad, s[k].grad), cos(p.grad, s[k].grad)) for k, p in net.named_parameters()]
            pairs += [("d_img", rel(leaves[0][0].grad, lv[0].grad), cos(leaves[0][0].grad, lv[0].grad)),
                      ("d_c", rel(leaves[0][1].grad, lv[1].grad), cos(leaves[0][1].grad, lv[1].grad))]
            rec[f"grads_vs_{tag}"] = summarise(pairs)
        out[f"D{which}_B{B}"] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    res = {}
    res["forward"] = section_forward()
    print(json.dumps(res["forward"], indent=1), flush=True)
    res["repro_3stage_B6"] = section_repro(3, 6)
    print(json.dumps(res["repro_3stage_B6"], indent=1), flush=True)
    res["step_1stage_B8"] = section_step(1, 8)
    res["step_3stage_B6"] = section_step(3, 6)
    if not args.quick:
        res["step_3stage_B24"] = section_step(3, 24)
        res["repro_3stage_B24"] = section_repro(3, 24)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "forward"}, indent=1))


if __name__ == "__main__":
    main()
