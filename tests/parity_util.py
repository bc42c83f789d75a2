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
