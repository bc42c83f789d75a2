"""Loss-curve tracking: N fused train steps (bf16 CUDA kernels) next to N oracle steps (fp32 torch on the same GPU,
TF32 off) from identical initial weights on identical per-step data / noise streams (north_star: "loss curves over
1,000 steps must track the reference within the stated tolerance").

GAN training is chaotic, so after a few hundred steps the two trajectories are different samples of the same process;
what is compared is (a) the per-step relative deviation over the first steps, (b) window means of every loss over the
whole run. Writes a small JSON / text summary (committed under profiles/).

    python tools/loss_curve.py --steps 1000 --branches 1 --batch 24 --out gpurun_out/loss_curve_1stage.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


PRINT_EVERY = int(os.environ.get("PRINT_EVERY", "100"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--branches", type=int, default=1)
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--window", type=int, default=100)
    ap.add_argument("--out", default="gpurun_out/loss_curve.json")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="arithmetic mode of the CUDA arm")
    ap.add_argument("--emu", action="store_true",
                    help="also run the bf16-EMULATING oracle on the same streams: its distance from the fp32 oracle is the "
                         "envelope an ideal bf16-storage implementation stays in")
    ap.add_argument("--rotate", type=int, default=0,
                    help="reuse this many batches round robin (bench.py rotates 3): a memorising discriminator saturates and "
                         "the run leaves the regime of healthy GAN training — checks that both arms fail the same way")
    ap.add_argument("--n-classes", type=int, default=6)
    args = ap.parse_args()
    from oracle.stackgan_oracle import Cfg, OracleTrainer, emulate_bf16
    from sg2b200 import config as _cfgmod
    _cfgmod.set_precision(args.precision)
    from sg2b200 import config, trainer, utils
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = config.cfg
    cfg.TREE.BRANCH_NUM = args.branches
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    netG, netsD = utils.build_networks(cfg, dev)
    ocfg = Cfg(BRANCH_NUM=args.branches)
    orc = OracleTrainer(ocfg, {k: v.detach().clone() for k, v in netG.state_dict().items()},
                        [{k: v.detach().clone() for k, v in d.state_dict().items()} for d in netsD], device=dev)
    orq = OracleTrainer(ocfg, {k: v.detach().clone() for k, v in netG.state_dict().items()},
                        [{k: v.detach().clone() for k, v in d.state_dict().items()} for d in netsD], device=dev) if args.emu else None
    tr = trainer.FusedTrainer(netG, netsD, cfg)
    nD = args.branches
    names = [f"errD{i}" for i in range(nD)] + ["errG_total", "kl", "cal"]
    ours, ref, emu = [], [], []
    for s in range(args.steps):
        b = utils.synthetic_batch(cfg, args.batch, seed=1000 + (s % args.rotate if args.rotate else s), device=dev,
                                  n_classes=args.n_classes)
        eps = torch.randn(args.batch, cfg.GAN.EMBEDDING_DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(s))
        lo = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=eps).cpu().tolist()
        o = orc.step(dict(z=b["z"], emb=b["emb"], eps=eps, real=b["real"], wrong=b["wrong"], labels=b["labels"].tolist()))
        lr = [float(x) for x in o["errD"]] + [float(o["errG_total"]), float(o["kl"]), float(o["cal"])]
        ours.append(lo)
        ref.append(lr)
        if orq is not None:
            with emulate_bf16():
                q = orq.step(dict(z=b["z"], emb=b["emb"], eps=eps, real=b["real"], wrong=b["wrong"], labels=b["labels"].tolist()))
            emu.append([float(x) for x in q["errD"]] + [float(q["errG_total"]), float(q["kl"]), float(q["cal"])])
        bad = not all(v == v for v in lo) or not all(v == v for v in lr)
        if s % PRINT_EVERY == 0 or bad or (not args.rotate and (max(lo[:nD]) > 4 or max(lr[:nD]) > 4)):
            print(f"step {s}: ours {[round(v, 4) for v in lo]} ref {[round(v, 4) for v in lr]}", flush=True)
    A, R = torch.tensor(ours, dtype=torch.float64), torch.tensor(ref, dtype=torch.float64)
    rel = (A - R).abs() / (R.abs() + 1e-3)
    out = {"steps": args.steps, "branches": args.branches, "batch": args.batch, "names": names,
           "first_steps_max_rel": {n: float(rel[:10, i].max()) for i, n in enumerate(names)},
           "first_50_mean_rel": {n: float(rel[:50, i].mean()) for i, n in enumerate(names)},
           "windows": []}
    W = args.window
    for w0 in range(0, args.steps, W):
        a, r = A[w0:w0 + W].mean(0), R[w0:w0 + W].mean(0)
        out["windows"].append({"steps": [w0, min(args.steps, w0 + W)],
                               "ours_mean": [round(float(v), 5) for v in a], "ref_mean": [round(float(v), 5) for v in r],
                               "rel_dev_of_means": [round(float(abs(x - y) / (abs(y) + 1e-3)), 4) for x, y in zip(a, r)]})
    out["whole_run_rel_dev_of_means"] = {n: float(abs(A[:, i].mean() - R[:, i].mean()) / (abs(R[:, i].mean()) + 1e-3))
                                         for i, n in enumerate(names)}
    out["finite"] = bool(torch.isfinite(A).all())
    out["precision"] = args.precision
    if emu:
        Q = torch.tensor(emu, dtype=torch.float64)
        relq = (Q - R).abs() / (R.abs() + 1e-3)
        out["envelope_bf16_emulating_oracle_vs_fp32_oracle"] = {
            "first_steps_max_rel": {n: float(relq[:10, i].max()) for i, n in enumerate(names)},
            "first_50_mean_rel": {n: float(relq[:50, i].mean()) for i, n in enumerate(names)},
            "whole_run_rel_dev_of_means": {n: float(abs(Q[:, i].mean() - R[:, i].mean()) / (abs(R[:, i].mean()) + 1e-3))
                                           for i, n in enumerate(names)}}
        # per 100-step window: our deviation from the fp32 oracle next to the ideal-bf16 envelope (mean |rel| of errG_total
        # and of the discriminator losses)
        gi = names.index("errG_total")
        wins = []
        for w0 in range(0, args.steps, W):
            sl = slice(w0, w0 + W)
            wins.append({"steps": [w0, min(args.steps, w0 + W)],
                         "ours_mean_rel": {"errG_total": float(rel[sl, gi].mean()), "errD": float(rel[sl, :nD].mean())},
                         "envelope_mean_rel": {"errG_total": float(relq[sl, gi].mean()), "errD": float(relq[sl, :nD].mean())}})
        out["windows_vs_envelope"] = wins
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("first_steps_max_rel", "first_50_mean_rel", "whole_run_rel_dev_of_means", "finite")}))


if __name__ == "__main__":
    main()
