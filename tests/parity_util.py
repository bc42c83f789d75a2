"""Shared helpers for the GPU parity tests: build sg2b200 networks and the oracle from the same state dicts."""
import torch

from oracle.stackgan_oracle import Cfg, init_d_state, init_g_state


def set_cfg(c: Cfg):
    """Point sg2b200's global cfg at an oracle Cfg."""
    from sg2b200.config import cfg
    cfg.TREE.BRANCH_NUM = c.BRANCH_NUM
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM = c.GF_DIM, c.DF_DIM
    cfg.GAN.EMBEDDING_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = c.EMBEDDING_DIM, c.Z_DIM, c.R_NUM
    cfg.TEXT.DIMENSION = c.TEXT_DIM
    return cfg


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def fp32_strict():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def make_g(c: Cfg, seed=0, device="cuda"):
    from sg2b200 import model
    set_cfg(c)
    torch.manual_seed(seed)
    sd = init_g_state(c)
    net = model.G_NET()
    net.load_state_dict(sd)
    return net.to(device), {k: v.clone().to(device) for k, v in sd.items()}


def make_d(c: Cfg, which, seed=0, device="cuda"):
    from sg2b200 import model
    set_cfg(c)
    torch.manual_seed(seed + 10 + which)
    sd = init_d_state(c, which)
    net = [model.D_NET64, model.D_NET128, model.D_NET256][which]()
    net.load_state_dict(sd)
    return net.to(device), {k: v.clone().to(device) for k, v in sd.items()}


def report(pairs, tol):
    """pairs: list of (name, rel_err). Returns (ok, message listing the worst offenders)."""
    worst = sorted(pairs, key=lambda t: -t[1])[:8]
    ok = all(e <= tol or e != e for _, e in pairs) and all(e == e for _, e in pairs)
    return ok, "worst rel errors: " + ", ".join(f"{n}={e:.3e}" for n, e in worst)


# ------------------------------------------------------------------------------------------------ train-step helpers
def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.detach().flatten().double(), b.detach().flatten().double(), dim=0))


def f64_state(sd):
    """State dict for a float64 run of the oracle (the ground truth of the reference's arithmetic)."""
    return {k: (v.detach().double() if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}


def f64_batch(b):
    cv = lambda t: t.double() if torch.is_tensor(t) and t.is_floating_point() else t
    return {k: ([cv(t) for t in v] if isinstance(v, list) else cv(v)) for k, v in b.items()}


def build_trainer_and_oracles(branches, seed=0, n_oracles=1, lr=None, oracle_f64=False):
    """sg2b200 networks + FusedTrainer and `n_oracles` OracleTrainers, all starting from the same weights.
    oracle_f64: the oracles run in float64."""
    from oracle.stackgan_oracle import OracleTrainer
    from sg2b200 import trainer, utils
    fp32_strict()
    ocfg = Cfg(BRANCH_NUM=branches)
    if lr is not None:
        ocfg.LR_G = ocfg.LR_D = lr
    cfg = set_cfg(ocfg)
    torch.manual_seed(seed)
    netG, netsD = utils.build_networks(cfg, "cuda")
    gs = {k: v.detach().clone() for k, v in netG.state_dict().items()}
    dss = [{k: v.detach().clone() for k, v in d.state_dict().items()} for d in netsD]
    if oracle_f64:
        gs, dss = f64_state(gs), [f64_state(d) for d in dss]
    orcs = [OracleTrainer(ocfg, gs, dss, device="cuda") for _ in range(n_oracles)]
    tr = trainer.FusedTrainer(netG, netsD, cfg, lr_g=lr, lr_d=lr)
    return cfg, ocfg, netG, netsD, tr, orcs


def train_batch(cfg, B, seed, n_classes=3):
    from sg2b200 import utils
    b = utils.synthetic_batch(cfg, B, seed=seed, device="cuda", n_classes=n_classes)
    b["eps"] = torch.randn(B, cfg.GAN.EMBEDDING_DIM, generator=torch.Generator().manual_seed(seed + 1)).cuda()
    return b


def oracle_step(orc, b, **kw):
    if next(iter(orc.g.values())).dtype == torch.float64:
        b = f64_batch(b)
    return orc.step(dict(z=b["z"], emb=b["emb"], eps=b["eps"], real=b["real"], wrong=b["wrong"],
                         labels=b["labels"].tolist()), **kw)


def loss_vector(o):
    """Oracle step output -> [errD_0.., errG_total, kl, cal] like FusedTrainer.step returns."""
    return [float(e) for e in o["errD"]] + [float(o["errG_total"]), float(o["kl"]), float(o["cal"])]


def bucket_grads(tr):
    """name -> gradient tensor (reference OIHW shapes) out of the fused trainer's flat gradient buckets."""
    out = {}
    for tag, net, b in [("G", tr.netG, tr.bG)] + [(f"D{i}", d, tr.bD[i]) for i, d in enumerate(tr.netsD)]:
        for k, p in net.named_parameters():
            out[f"{tag}.{k}"] = b.views[p]
    return out


def oracle_grads(o):
    out = {f"G.{k}": v for k, v in o["grads_g"].items()}
    for i, gd in enumerate(o["grads_d"]):
        out.update({f"D{i}.{k}": v for k, v in gd.items()})
    return out


def snapshot_diff(a, b):
    """Largest relative difference between two FusedTrainer.snapshot()s, and whether they are bit-identical."""
    worst, same = 0.0, True
    for ba, bb in zip(a["buckets"], b["buckets"]):
        for k in ba:
            same = same and torch.equal(ba[k], bb[k])
            if ba[k].is_floating_point() and ba[k].numel() > 0:
                worst = max(worst, rel(ba[k], bb[k]))
    for la, lb in zip(a["buffers"], b["buffers"]):
        for ta, tb in zip(la, lb):
            same = same and torch.equal(ta, tb)
            if ta.is_floating_point():
                worst = max(worst, rel(ta, tb))
    return worst, same
