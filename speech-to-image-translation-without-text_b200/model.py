"""Drop-in replacements for the reference's StackGAN_v2/model.py module API.

    G_NET()(z_code, text_embedding)  -> ([img64, img128, img256][:BRANCH_NUM], mu, logvar)   model.py:327-354
    D_NET64/128/256()(x_var, c_code) -> ([cond (B,), uncond (B,)], x_immediate (B, 8192))    model.py:424-445

Constructor (no arguments, reads the global cfg), parameter names / shapes / ORDER, buffer names, class names that
`weights_init` dispatches on ('Conv', 'BatchNorm', 'Linear'; trainer.py:65-75), `.cuda()`, `.train()/.eval()`,
`state_dict()` (checkpoint-compatible, DataParallel-wrappable) all match the reference. The nn.Conv2d / BatchNorm /
Linear objects are parameter holders only: forward and backward run on the sg2b200 CUDA kernels (nets.py) through
one autograd Function per network. There is no eager/CPU fallback.
"""
import torch
import torch.nn as nn

from . import nets
from . import config as _config
from .config import active_cfg

__all__ = ["G_NET", "D_NET64", "D_NET128", "D_NET256", "D_NET512", "D_NET1024", "INCEPTION_V3", "GLU"]


class GLU(nn.Module):
    """Marker only (no parameters): the gate is fused into the BatchNorm-apply kernel."""

    def forward(self, x):
        raise RuntimeError("sg2b200 containers are parameter holders; call the top-level network")


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices equal to the reference's."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


def _conv3(cin, cout):
    return nn.Conv2d(cin, cout, 3, 1, 1, bias=False)


def _up(cin, cout):       # indices: 0 upsample, 1 conv, 2 BN, 3 GLU   (model.py:133-140)
    return nn.Sequential(_Slot("nearest2x (fused into the conv)"), _conv3(cin, cout * 2), nn.BatchNorm2d(cout * 2), GLU())


def _c3_glu(cin, cout):   # model.py:144-150
    return nn.Sequential(_conv3(cin, cout * 2), nn.BatchNorm2d(cout * 2), GLU())


def _c3_lrelu(cin, cout):  # model.py:358-365
    return nn.Sequential(_conv3(cin, cout), nn.BatchNorm2d(cout), _Slot("LeakyReLU(0.2)"))


def _down(cin, cout):     # model.py:369-376
    return nn.Sequential(nn.Conv2d(cin, cout, 4, 2, 1, bias=False), nn.BatchNorm2d(cout), _Slot("LeakyReLU(0.2)"))


class ResBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.block = nn.Sequential(_conv3(c, c * 2), nn.BatchNorm2d(c * 2), GLU(), _conv3(c, c), nn.BatchNorm2d(c))


class CA_NET(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fc = nn.Linear(cfg.TEXT.DIMENSION, cfg.GAN.EMBEDDING_DIM * 4, bias=True)
        self.relu = GLU()


class INIT_STAGE_G(nn.Module):
    def __init__(self, ngf, cfg):
        super().__init__()
        in_dim = cfg.GAN.Z_DIM + cfg.GAN.EMBEDDING_DIM
        self.fc = nn.Sequential(nn.Linear(in_dim, ngf * 4 * 4 * 2, bias=False), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self.upsample1 = _up(ngf, ngf // 2)
        self.upsample2 = _up(ngf // 2, ngf // 4)
        self.upsample3 = _up(ngf // 4, ngf // 8)
        self.upsample4 = _up(ngf // 8, ngf // 16)


class NEXT_STAGE_G(nn.Module):
    def __init__(self, ngf, cfg):
        super().__init__()
        self.jointConv = _c3_glu(ngf + cfg.GAN.EMBEDDING_DIM, ngf)
        self.residual = nn.Sequential(*[ResBlock(ngf) for _ in range(cfg.GAN.R_NUM)])
        self.upsample = _up(ngf, ngf // 2)


class GET_IMAGE_G(nn.Module):
    def __init__(self, ngf):
        super().__init__()
        self.img = nn.Sequential(_conv3(ngf, 3), _Slot("Tanh"))


def _check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sg2b200 networks run on CUDA (sm_100a) only; there is no CPU fallback")


class _GFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, training, z, emb, eps, *params):
        imgs, mu, logvar, tape = engine.forward(z.detach().contiguous().float(), emb.detach().contiguous().float(),
                                                eps, training)
        ctx.engine, ctx.tape, ctx.n = engine, tape, len(imgs)
        ctx.params = params
        return (*imgs, mu, logvar)

    @staticmethod
    def backward(ctx, *douts):
        n = ctx.n
        cg = lambda t: None if t is None else t.contiguous().float()
        dimgs = [cg(d) for d in douts[:n]]
        grads = ctx.engine.backward(ctx.tape, dimgs, cg(douts[n]), cg(douts[n + 1]))
        ctx.tape = None
        return (None, None, None, None, None) + tuple(grads.get(p) for p in ctx.params)



class G_NET(nn.Module):
    def __init__(self):
        super().__init__()
        cfg = active_cfg()
        if not cfg.GAN.B_CONDITION:
            raise NotImplementedError("only the conditional path (B_CONDITION: True) is live in the reference")
        if not 1 <= cfg.TREE.BRANCH_NUM <= 3:
            raise NotImplementedError("BRANCH_NUM > 3 is never reached by any reference cfg")
        self.gf_dim = cfg.GAN.GF_DIM
        self.ef_dim = cfg.GAN.EMBEDDING_DIM
        self.ca_net = CA_NET(cfg)
        self.h_net1 = INIT_STAGE_G(self.gf_dim * 16, cfg)
        self.img_net1 = GET_IMAGE_G(self.gf_dim)
        if cfg.TREE.BRANCH_NUM > 1:
            self.h_net2 = NEXT_STAGE_G(self.gf_dim, cfg)
            self.img_net2 = GET_IMAGE_G(self.gf_dim // 2)
        if cfg.TREE.BRANCH_NUM > 2:
            self.h_net3 = NEXT_STAGE_G(self.gf_dim // 2, cfg)
            self.img_net3 = GET_IMAGE_G(self.gf_dim // 4)
        self._cfg = cfg
        self._engine = None
        self._precision = None

    def engine(self):
        if self._engine is None:
            self._engine = nets.GEngine(self, self._cfg, precise=(self._precision or _config.PRECISION) == "fp32")
        return self._engine

    def set_precision(self, p):
        """'bf16' | 'fp32' (see config.PRECISION): rebuilds the execution engine in that arithmetic mode."""
        if p not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._precision, self._engine = p, None

    def forward(self, z_code, text_embedding=None, eps=None):
        if text_embedding is None:
            raise NotImplementedError("unconditional G_NET is dead code in the reference")
        _check_cuda(z_code, text_embedding)
        if eps is None:
            # same draw as CA_NET.reparametrize (model.py:190-193): default CUDA generator, N(0,1), (B, EMBEDDING_DIM)
            eps = torch.empty(z_code.shape[0], self.ef_dim, device=z_code.device, dtype=torch.float32).normal_()
        params = tuple(self.parameters())
        out = _GFn.apply(self.engine(), self.training, z_code, text_embedding, eps, *params)
        return list(out[:-2]), out[-2], out[-1]


class _DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, training, need_dimg, need_w, img, c, *params):
        cond, uncond, x_imm, tape = engine.forward(img.detach().contiguous().float(), c.detach().contiguous().float(),
                                                   training)
        ctx.engine, ctx.tape, ctx.params = engine, tape, params
        ctx.need_dimg, ctx.need_w = need_dimg, need_w
        return cond, uncond, x_imm

    @staticmethod
    def backward(ctx, dcond, duncond, dx_imm):
        need_dc = ctx.needs_input_grad[5]
        cg = lambda t: None if t is None else t.contiguous().float()
        grads, dimg, dc = ctx.engine.backward(ctx.tape, cg(dcond), cg(duncond), cg(dx_imm),
                                              ctx.need_dimg and ctx.needs_input_grad[4], need_dc, ctx.need_w)
        ctx.tape = None
        return (None, None, None, None, dimg, dc) + tuple(grads.get(p) for p in ctx.params)


class _DBase(nn.Module):
    """Shared machinery of D_NET64/128/256. `skip_param_grads` (context manager) drops the weight gradients of a
    pass whose D gradients the caller discards (the G step: trainer.py:385 zeroes them before use)."""
    data_image_grads = False   # dgrad into leaf (dataset) images is skipped: nothing reads real_imgs.grad

    def __init__(self):
        super().__init__()
        cfg = active_cfg()
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = cfg.GAN.EMBEDDING_DIM
        self._cfg = cfg
        self._engine = None
        self._precision = None
        self._need_w = True
        ndf = self.df_dim
        self.img_code_s16 = nn.Sequential(      # indices as in model.py:380-398
            nn.Conv2d(3, ndf, 4, 2, 1, bias=False), _Slot("LeakyReLU(0.2)"),
            nn.Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), _Slot("LeakyReLU(0.2)"),
            nn.Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 4), _Slot("LeakyReLU(0.2)"),
            nn.Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 8), _Slot("LeakyReLU(0.2)"))

    def _tail(self):
        ndf, efg = self.df_dim, self.ef_dim
        self.logits = nn.Sequential(nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=4), _Slot("Sigmoid"))
        self.jointConv = _c3_lrelu(ndf * 8 + efg, ndf * 8)
        self.uncond_logits = nn.Sequential(nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=4), _Slot("Sigmoid"))

    def engine(self):
        if self._engine is None:
            self._engine = nets.DEngine(self, self._cfg, precise=(self._precision or _config.PRECISION) == "fp32")
        return self._engine

    set_precision = G_NET.set_precision

    def skip_param_grads(self):
        net = self

        class _Ctx:
            def __enter__(self):
                net._need_w = False

            def __exit__(self, *a):
                net._need_w = True

        return _Ctx()

    def forward(self, x_var, c_code=None):
        if c_code is None:
            raise NotImplementedError("unconditional D is dead code in the reference")
        _check_cuda(x_var, c_code)
        need_dimg = x_var.requires_grad and (x_var.grad_fn is not None or self.data_image_grads)
        params = tuple(self.parameters())
        cond, uncond, x_imm = _DFn.apply(self.engine(), self.training, need_dimg, self._need_w, x_var, c_code, *params)
        return [cond, uncond], x_imm


class D_NET64(_DBase):
    def __init__(self):
        super().__init__()
        self._tail()


class D_NET128(_DBase):
    def __init__(self):
        super().__init__()
        ndf = self.df_dim
        self.img_code_s32 = _down(ndf * 8, ndf * 16)
        self.img_code_s32_1 = _c3_lrelu(ndf * 16, ndf * 8)
        self._tail()


class D_NET256(_DBase):
    def __init__(self):
        super().__init__()
        ndf = self.df_dim
        self.img_code_s32 = _down(ndf * 8, ndf * 16)
        self.img_code_s64 = _down(ndf * 16, ndf * 32)
        self.img_code_s64_1 = _c3_lrelu(ndf * 32, ndf * 16)
        self.img_code_s64_2 = _c3_lrelu(ndf * 16, ndf * 8)
        self._tail()


class _OutOfScope(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(
            f"{type(self).__name__} is outside the accelerated hot path (never built for BRANCH_NUM <= 3 / needs "
            "downloaded Inception weights); import it from the reference's model.py if required")


class D_NET512(_OutOfScope):
    pass


class D_NET1024(_OutOfScope):
    pass


class INCEPTION_V3(_OutOfScope):
    pass
