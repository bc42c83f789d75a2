"""Host-side helpers shared by bench.py, the tests and user code."""
import torch
import torch.nn as nn


def weights_init(m):
    """Same initialisation the reference applies with net.apply(weights_init) (trainer.py:65-75):
    orthogonal (gain 1) for Conv / Linear weights, BN weight ~ N(1, 0.02), BN / Linear bias 0."""
    name = m.__class__.__name__
    if name.find("Conv") != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
    elif name.find("BatchNorm") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif name.find("Linear") != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def build_networks(cfg, device="cuda"):
    """G_NET + the BRANCH_NUM discriminators, initialised like trainer.load_network (trainer.py:162-184) does."""
    from . import model
    netG = model.G_NET().to(device)
    netG.apply(weights_init)
    netsD = [cls().to(device) for cls in (model.D_NET64, model.D_NET128, model.D_NET256)[:cfg.TREE.BRANCH_NUM]]
    for d in netsD:
        d.apply(weights_init)
    return netG, netsD


def install_as_reference_model():
    """Make `from model import G_NET, D_NET64, ...` (trainer.py:24) resolve to sg2b200.model, so the reference's
    unmodified main.py / trainer.py run on the CUDA kernels. Call before importing the reference's trainer."""
    import sys
    from . import model
    sys.modules["model"] = model
    return model


def synthetic_batch(cfg, batch, seed, device="cpu", pin=False, n_classes=200):
    """Synthetic inputs of the benchmark: z, speech embedding, real / wrong image pyramids
    U(-1, 1) fp32 NCHW, class labels."""
    g = torch.Generator().manual_seed(seed)
    out = {"z": torch.randn(batch, cfg.GAN.Z_DIM, generator=g),
           "emb": torch.randn(batch, cfg.TEXT.DIMENSION, generator=g)}
    out["real"], out["wrong"] = [], []
    for i in range(cfg.TREE.BRANCH_NUM):
        s = 64 * 2 ** i
        out["real"].append(torch.rand(batch, 3, s, s, generator=g) * 2 - 1)
        out["wrong"].append(torch.rand(batch, 3, s, s, generator=g) * 2 - 1)
    lab = torch.randint(0, n_classes, (batch,), generator=g, dtype=torch.int32)
    if batch > 1:
        lab[1] = lab[0]          # at least one same-class pair so the class-aware loss is active
    out["labels"] = lab

    def mv(t):
        if pin:
            t = t.pin_memory()
        return t.to(device) if device != "cpu" else t
    return {k: ([mv(t) for t in v] if isinstance(v, list) else mv(v)) for k, v in out.items()}
