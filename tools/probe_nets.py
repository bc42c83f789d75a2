"""Diagnostic (not pytest): per-tensor error table of G_NET / D_NET* vs the oracle, plus first timings."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.stackgan_oracle import Cfg, d_forward, g_forward, is_param, emulate_bf16
from tests.parity_util import fp32_strict, make_d, make_g, rel


def g_probe(B, branches):
    cfg = Cfg(BRANCH_NUM=branches)
    fp32_strict()
    net, sd = make_g(cfg, seed=1)
    for k in sd:
        if is_param(k):
            sd[k].requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, cfg.Z_DIM, generator=g).cuda()
    emb = torch.randn(B, cfg.TEXT_DIM, generator=g).cuda()
    eps = torch.randn(B, cfg.EMBEDDING_DIM, generator=g).cuda()
    imgs, mu, logvar = net(z, emb, eps=eps)
    torch.cuda.synchronize()
    import copy
    sdq = {k: v.detach().clone().requires_grad_(is_param(k)) if v.is_floating_point() else v.clone() for k, v in sd.items()}
    oimgs, omu, ologvar = g_forward(sd, z, emb, eps, cfg, True)
    with emulate_bf16():
        qimgs, qmu, qlogvar = g_forward(sdq, z, emb, eps, cfg, True)
    print(f"G fwd B={B} branches={branches} vs fp32:", [f"{rel(a, b):.3e}" for a, b in zip(imgs, oimgs)], f"mu {rel(mu, omu):.2e} logvar {rel(logvar, ologvar):.2e}")
    print(f"                              vs bf16-emu:", [f"{rel(a, b):.3e}" for a, b in zip(imgs, qimgs)], " (emu vs fp32:", [f"{rel(a, b):.3e}" for a, b in zip(qimgs, oimgs)], ")")
    rs = [torch.randn(i.shape, generator=g).cuda() for i in oimgs]
    rmu, rlv = torch.randn(mu.shape, generator=g).cuda(), torch.randn(mu.shape, generator=g).cuda()
    L = lambda im, m, lv: sum((a * r).sum() for a, r in zip(im, rs)) + (m * rmu).sum() + (lv * rlv).sum()
    L(imgs, mu, logvar).backward(); L(oimgs, omu, ologvar).backward(); L(qimgs, qmu, qlogvar).backward()
    torch.cuda.synchronize()
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))
    errs = [(k, rel(p.grad, sd[k].grad), rel(p.grad, sdq[k].grad), cos(p.grad, sd[k].grad)) for k, p in net.named_parameters()]
    for k, e, eq, cs in errs:
        print(f"   grad {k:45s} vs fp32 {e:.3e}  vs emu {eq:.3e}  cos(fp32) {cs:.5f}")


def d_probe(which, B):
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
    img = base.clone().requires_grad_(True); c = c0.clone().requires_grad_(True)
    oimg = base.clone().requires_grad_(True); oc = c0.clone().requires_grad_(True)
    (cond, uncond), x_imm = net(img * 1.0, c * 1.0)
    sdq = {k: v.detach().clone().requires_grad_(is_param(k)) if v.is_floating_point() else v.clone() for k, v in sd.items()}
    qimg = base.clone().requires_grad_(True); qc = c0.clone().requires_grad_(True)
    (ocond, ouncond), ox = d_forward(sd, oimg * 1.0, oc * 1.0, which, cfg, True)
    with emulate_bf16():
        (qcond, quncond), qx = d_forward(sdq, qimg * 1.0, qc * 1.0, which, cfg, True)
    print(f"D{which} fwd B={B} vs fp32: cond {rel(cond, ocond):.3e} uncond {rel(uncond, ouncond):.3e} x_imm {rel(x_imm, ox):.3e}")
    print(f"              vs emu : cond {rel(cond, qcond):.3e} uncond {rel(uncond, quncond):.3e} x_imm {rel(x_imm, qx):.3e}")
    r1, r2 = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
    r3 = torch.randn(ox.shape, generator=g).cuda() * 0.01
    ((cond * r1).sum() + (uncond * r2).sum() + (x_imm * r3).sum()).backward()
    ((ocond * r1).sum() + (ouncond * r2).sum() + (ox * r3).sum()).backward()
    ((qcond * r1).sum() + (quncond * r2).sum() + (qx * r3).sum()).backward()
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))
    errs = [(k, rel(p.grad, sd[k].grad), rel(p.grad, sdq[k].grad), cos(p.grad, sd[k].grad)) for k, p in net.named_parameters()]
    errs += [("d_img", rel(img.grad, oimg.grad), rel(img.grad, qimg.grad), cos(img.grad, oimg.grad)), ("d_c", rel(c.grad, oc.grad), rel(c.grad, qc.grad), cos(c.grad, oc.grad))]
    for k, e, eq, cs in errs:
        print(f"   grad {k:45s} vs fp32 {e:.3e}  vs emu {eq:.3e}  cos(fp32) {cs:.5f}")


def timing(B=24):
    """First wall/device timings of the module API at the production shape (eager launches, no graph)."""
    cfg = Cfg()
    net, _ = make_g(cfg, seed=1)
    ds = [make_d(cfg, i)[0] for i in range(3)]
    z = torch.randn(B, cfg.Z_DIM).cuda(); emb = torch.randn(B, cfg.TEXT_DIM).cuda()
    def g_fb():
        imgs, mu, lv = net(z, emb)
        (sum(i.mean() for i in imgs) + mu.mean() + lv.mean()).backward()
        return imgs, mu
    def run(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.time(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.time() - t0) * 1000 / n
    print("G fwd+bwd B=%d: dev %.2f ms wall %.2f ms" % ((B,) + run(g_fb)))
    with torch.no_grad():
        print("G fwd only (no_grad): dev %.2f ms wall %.2f ms" % run(lambda: net(z, emb)))
    imgs, mu = g_fb()
    for i, d in enumerate(ds):
        im = imgs[i].detach()
        def d_fb():
            (cnd, unc), x = d(im, mu.detach())
            (cnd.mean() + unc.mean()).backward()
        print("D%d fwd+bwd (weights only) B=%d: dev %.2f ms wall %.2f ms" % ((i, B) + run(d_fb)))


if __name__ == "__main__":
    for fn, args in [(g_probe, (8, 1)), (g_probe, (4, 3)), (d_probe, (0, 8)), (d_probe, (1, 6)), (d_probe, (2, 4)), (timing, ())]:
        try:
            fn(*args)
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()
