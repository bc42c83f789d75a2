"""GPU unit parity of each kernel family against plain torch fp32 on the SAME (bf16-rounded) inputs, called
through the C ABI (sg2b200.ops -> ctypes -> libsg2b200.so). Tolerances are stated per test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bf(t):
    return t.bfloat16().float()


@pytest.fixture(autouse=True)
def _strict():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


# --------------------------------------------------------------------------------------------- convolutions
CONV_CASES = [  # kind, B, H, W, Cin, Cout
    (0, 2, 16, 16, 64, 64), (0, 3, 8, 8, 32, 32), (0, 2, 32, 32, 160, 64), (0, 2, 16, 16, 64, 192),
    (0, 5, 4, 4, 128, 256), (0, 2, 32, 32, 16, 32), (1, 3, 4, 4, 64, 64), (1, 2, 16, 16, 32, 32),
    (2, 2, 16, 16, 64, 128), (2, 5, 8, 8, 128, 256), (3, 1, 1, 640, 64, 64),
    # tile-resident kernel (output grid >= 16 x 8): partial edge tiles, four parity-plane sources, every swizzle width
    (0, 1, 24, 20, 32, 96), (2, 2, 32, 32, 64, 128), (2, 1, 48, 40, 32, 64), (1, 1, 24, 20, 64, 32),
    (0, 2, 32, 16, 128, 256), (2, 1, 32, 32, 16, 32),
    # CTA-pair gather kernel (256-channel tiles, >= 2 pixel tiles): odd tile count (phantom peer), parity groups, N > 256
    (2, 12, 16, 16, 128, 256), (0, 20, 4, 4, 256, 512), (1, 24, 4, 4, 256, 256),
]


@pytest.mark.parametrize("kind,B,H,W,Ci,Co", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(kind, B, H, W, Ci, Co):
    """tcgen05 implicit GEMM vs F.conv2d (+interpolate / stride): bf16 operands, fp32 accumulation; rel err <= 5e-3
    (the fused-upsample kind pre-sums taps before rounding to bf16, which moves individual weights by <= 1 ulp)."""
    from sg2b200 import ops
    from tools.probe_conv import pack_ref, ref_fwd, unpack_wgrad_ref
    g = torch.Generator().manual_seed(kind * 100 + Ci + Co)
    k = {0: 3, 1: 3, 2: 4, 3: 1}[kind]
    x = _bf(torch.randn(B, Ci, H, W, generator=g)).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).cuda().requires_grad_(True)
    y_ref = ref_fwd(kind, x, _bf(w) if kind != 1 else w)
    dy = _bf(torch.randn(y_ref.shape, generator=g)).cuda()
    dx_ref, dw_ref = torch.autograd.grad(y_ref, (x, w), dy)
    s1, s2 = ops.pack_shapes(kind, Co, Ci)
    wpk = torch.empty(s1, device="cuda", dtype=torch.bfloat16)
    wpkT = torch.empty(s2, device="cuda", dtype=torch.bfloat16)
    ops.pack_weights(kind, w.detach().contiguous(), wpk, wpkT, Co, Ci, Co, Ci)
    rk, rkT = pack_ref(kind, w.detach())
    assert torch.equal(wpk, rk.bfloat16()) and torch.equal(wpkT, rkT.bfloat16())     # packing is bit-exact
    # the fused trainer stores masters [Cout][kh][kw][Cin]: same packs from that layout, and the dgrad operand as a
    # per-tap transpose of the bf16 fprop operand
    w_ohwi = w.detach().permute(0, 2, 3, 1).contiguous()
    wpk2, wpkT2 = torch.empty_like(wpk), torch.empty_like(wpkT)
    ops.pack_weights(kind, w_ohwi, wpk2, wpkT2, Co, Ci, Co, Ci, ohwi=True)
    assert torch.equal(wpk2, wpk) and torch.equal(wpkT2, wpkT)
    if kind != 1 and Co % 8 == 0 and Ci % 8 == 0:
        wpkT3 = torch.full_like(wpkT, float("nan"))
        ops.pack_transpose(kind, wpk, wpkT3, Co, Ci)
        assert torch.equal(wpkT3, wpkT)
    xn = x.detach().permute(0, 2, 3, 1).contiguous().bfloat16()
    dyn = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    y = ops.conv_fprop(kind, xn, wpk, Co, splitk=1)
    assert _rel(y.float().permute(0, 3, 1, 2), y_ref) < 5e-3
    st = torch.zeros(2 * Co, device="cuda", dtype=torch.float64)                                         # BN statistics from the epilogue
    y3, fused = ops.conv_fprop(kind, xn, wpk, Co, splitk=1, stats=st)
    assert fused and torch.equal(y3, y)
    yf = y.float().reshape(-1, Co)
    assert _rel(st[:Co], yf.sum(0)) < 1e-4 and _rel(st[Co:], (yf * yf).sum(0)) < 1e-5
    y2 = ops.conv_fprop(kind, xn, wpk, Co, splitk=3)                                   # split-K: ordered slab sum
    assert _rel(y2.float().permute(0, 3, 1, 2), y_ref) < 5e-3
    dx = ops.conv_dgrad(kind, dyn, wpkT, B, H, W, Ci, splitk=1)
    assert _rel(dx.float().permute(0, 3, 1, 2), dx_ref) < 5e-3
    dwpk = torch.zeros(Co, ops.JOBS[kind], Ci, device="cuda")
    ops.conv_wgrad(kind, xn, dyn, dwpk)
    if ops.DETERMINISTIC:
        # first=True writes every element (no clearing needed) and the ordered slab reduction is reproducible bit for bit
        dwpk_b = torch.full_like(dwpk, float("nan"))
        ops.conv_wgrad(kind, xn, dyn, dwpk_b, first=True)
        assert torch.equal(dwpk_b, dwpk)
        ops.conv_wgrad(kind, xn, dyn, dwpk_b)                                          # second pass accumulates
        assert torch.equal(dwpk_b, dwpk + dwpk)
        assert torch.equal(ops.conv_fprop(kind, xn, wpk, Co, splitk=3), y2)
        for sk in (2, 3, 5):        # BatchNorm statistics out of the split-K reduction (cluster or slab path)
            st_s = torch.zeros(2 * Co, device="cuda", dtype=torch.float64)
            ys, _ = ops.conv_fprop(kind, xn, wpk, Co, splitk=sk, stats=st_s)
            ysf = ys.float().reshape(-1, Co)
            assert _rel(ys.float().permute(0, 3, 1, 2), y_ref) < 5e-3
            assert _rel(st_s[:Co], ysf.sum(0)) < 1e-4 and _rel(st_s[Co:], (ysf * ysf).sum(0)) < 1e-5, sk
    gk = torch.empty_like(dw_ref)
    ops.unpack_wgrad(kind, dwpk, gk, Co, Ci, Co, Ci, False)
    assert torch.allclose(gk, unpack_wgrad_ref(kind, dwpk, Co, Ci), atol=1e-5)          # unpack is a pure permute/sum
    gk2 = torch.empty(Co, k, k, Ci, device="cuda")
    ops.unpack_wgrad(kind, dwpk, gk2, Co, Ci, Co, Ci, False, ohwi=True)
    assert torch.equal(gk2.permute(0, 3, 1, 2), gk)
    assert _rel(gk, dw_ref) < 5e-3


# --------------------------------------------------------------------------------------------- BN + activations
@pytest.mark.parametrize("act", [0, 1, 2])
# (70000, 64) is > 4 MB with a ragged last tile: the cp.async.bulk-staged variants; the others use register loads
@pytest.mark.parametrize("P,C", [(24 * 64, 128), (1000, 32), (24, 4096), (96, 640), (70000, 64)])
def test_bn_act_forward_backward(act, P, C):
    """BatchNorm(train) + {identity(+residual), GLU, LeakyReLU(0.2)} vs torch; outputs are bf16 -> rel err <= 4e-3;
    statistics and dgamma/dbeta are fp32 -> <= 1e-4 / 2e-3."""
    from sg2b200 import ops
    g = torch.Generator().manual_seed(P + C + act)
    x = _bf(torch.randn(P, C, generator=g) * 1.7 + 0.3).cuda()
    gamma = (torch.randn(C, generator=g) * 0.1 + 1).cuda()
    beta = (torch.randn(C, generator=g) * 0.1).cuda()
    Co = C // 2 if act == 1 else C
    res = _bf(torch.randn(P, Co, generator=g)).cuda() if act == 0 else None
    dout = _bf(torch.randn(P, Co, generator=g)).cuda()
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    xb = x.bfloat16()
    st = ops.bn_stats32(C, xb.device)
    ops.bn_stats(xb, st)
    out, mean, rstd = ops.bn_act_fwd(xb, gamma, beta, act, None if res is None else res.bfloat16(), stats=st,
                                     running=(rm, rv, nbt))
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    z = F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    assert _rel(mean, x.mean(0)) < 1e-4 and _rel(rstd, 1 / torch.sqrt(x.var(0, unbiased=False) + 1e-5)) < 1e-4
    assert _rel(rm, rm2) < 1e-4 and _rel(rv, rv2) < 1e-4 and int(nbt) == 1
    if act == 1:
        o_ref = z[:, :Co] * torch.sigmoid(z[:, Co:])
    elif act == 2:
        o_ref = F.leaky_relu(z, 0.2)
    else:
        o_ref = z + res
    assert _rel(out.float(), o_ref) < 4e-3
    out_eval = ops.bn_act_fwd(xb, gamma, beta, act, None if res is None else res.bfloat16(), mean=mean.contiguous(),
                              rstd=rstd.contiguous())
    assert torch.equal(out_eval, out)          # eval-style call with given mean/rstd is the same arithmetic
    dx_ref, dg_ref, db_ref = torch.autograd.grad(o_ref, (xr, gr, br), dout)
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx = ops.bn_act_bwd(xb, dout.bfloat16(), mean, rstd, gamma, beta, act, dg, db, False)
    assert _rel(dx.float(), dx_ref) < 6e-3
    assert _rel(dg, dg_ref) < 2e-3 and _rel(db, db_ref) < 2e-3
    # accumulate mode adds on top
    ops.bn_act_bwd(xb, dout.bfloat16(), mean, rstd, gamma, beta, act, dg, db, True)
    assert _rel(dg, 2 * dg_ref) < 2e-3


def test_glu_backward_extreme_gate():
    """A sharply peaked gate channel with a few outlier pixels: after BatchNorm the outliers sit hundreds of standard
    deviations below zero (|z| ~ sqrt(pixels)), exp(-z) overflows fp32 and a backward pass that forms 1 - sigmoid as
    exp(-z) * sigmoid returned NaN for the whole channel (rotating-batch run, step 189: every generator gradient NaN
    while the fp32 oracle stayed finite). Forward and backward must stay finite and match torch."""
    from sg2b200 import ops
    P, C = 4096 * 16, 64
    g = torch.Generator().manual_seed(7)
    x = torch.randn(P, C, generator=g) * 1e-3
    x[:3, C // 2 + 5] = -40.0            # gate half, channel 5: z ~ -150 after normalisation (2^216 as exp2 argument)
    x[:3, C // 2 + 9] = 40.0             # and the saturated-open side
    x = _bf(x).cuda()
    gamma, beta = torch.ones(C).cuda(), torch.zeros(C).cuda()
    dout = _bf(torch.randn(P, C // 2, generator=g)).cuda()
    xb = x.bfloat16()
    st = ops.bn_stats32(C, xb.device)
    ops.bn_stats(xb, st)
    out, mean, rstd = ops.bn_act_fwd(xb, gamma, beta, 1, None, stats=st)
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx = ops.bn_act_bwd(xb, dout.bfloat16(), mean, rstd, gamma, beta, 1, dg, db, False)
    assert torch.isfinite(out.float()).all() and torch.isfinite(dx.float()).all()
    assert torch.isfinite(dg).all() and torch.isfinite(db).all()
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    o_ref = z[:, :C // 2] * torch.sigmoid(z[:, C // 2:])
    dx_ref, dg_ref, db_ref = torch.autograd.grad(o_ref, (xr, gr, br), dout)
    assert _rel(out.float(), o_ref) < 4e-3
    assert _rel(dx.float(), dx_ref) < 6e-3
    assert _rel(dg, dg_ref) < 2e-3 and _rel(db, db_ref) < 2e-3


def test_concat_c_and_backward():
    from sg2b200 import ops
    g = torch.Generator().manual_seed(1)
    B, H, W, E, Ch = 3, 8, 8, 128, 32
    c = torch.randn(B, E, generator=g).cuda()
    h = torch.randn(B, H, W, Ch, generator=g).cuda().bfloat16()
    cat = ops.concat_c(c, h)
    ref = torch.cat((c.bfloat16()[:, None, None, :].expand(B, H, W, E), h), -1)
    assert torch.equal(cat, ref)
    dcat = torch.randn(B, H, W, E + Ch, generator=g).cuda().bfloat16()
    dc = torch.zeros(B, E, device="cuda")
    dh = ops.concat_c_bwd(dcat, E, dc)
    assert torch.equal(dh, dcat[..., E:].contiguous())
    assert _rel(dc, dcat[..., :E].float().sum((1, 2))) < 1e-5


@pytest.mark.parametrize("kind,B,H,W,Ci,Co", [(0, 2, 32, 32, 32, 64), (2, 2, 32, 32, 64, 128), (0, 3, 4, 4, 128, 256),
                                                  (0, 20, 4, 4, 256, 512)])
def test_conv_dgrad_epilogue_operand(kind, B, H, W, Ci, Co):
    """dgrad with a residual gradient added / a LeakyReLU mask applied in the epilogue (or, for shapes on the gather
    kernel, by the separate kernel) equals dgrad followed by the separate kernel up to one bf16 rounding (2^-8)."""
    from sg2b200 import ops
    g = torch.Generator().manual_seed(5)
    k = 3 if kind == 0 else 4
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).cuda()
    s1, s2 = ops.pack_shapes(kind, Co, Ci)
    wpk = torch.empty(s1, device="cuda", dtype=torch.bfloat16)
    wpkT = torch.empty(s2, device="cuda", dtype=torch.bfloat16)
    ops.pack_weights(kind, w, wpk, wpkT, Co, Ci, Co, Ci)
    Ho, Wo = (H, W) if kind == 0 else (H // 2, W // 2)
    dy = torch.randn(B, Ho, Wo, Co, generator=g).cuda().bfloat16()
    src = torch.randn(B, H, W, Ci, generator=g).cuda().bfloat16()
    base = ops.conv_dgrad(kind, dy, wpkT, B, H, W, Ci).float()
    n0 = ops.launches()
    fused_add = ops.conv_dgrad(kind, dy, wpkT, B, H, W, Ci, epi=(src, ops.EPI_ADD)).float()
    n1 = ops.launches()
    fused_mask = ops.conv_dgrad(kind, dy, wpkT, B, H, W, Ci, epi=(src, ops.EPI_LRELU_MASK)).float()
    # output grids >= 16 x 8 run on the tile-resident kernel and the split 4 x 4 map on the cluster split-K kernel, both
    # with the operand in the epilogue: ONE launch (the slab path, SG2_CLUSTER_SPLITK=0, needs conv + finish)
    assert (n1 - n0 == 1) if (H >= 16 or ops.CLUSTER_SPLITK >= 2) else (n1 - n0 >= 2), n1 - n0
    ref_add = base + src.float()
    ref_mask = torch.where(src.float() > 0, base, 0.2 * base)
    scale = base.abs() + src.float().abs() + 1e-2           # the sum may cancel: error relative to the operands
    assert float(((fused_add - ref_add).abs() / scale).max()) < 2.0 ** -7
    assert float(((fused_mask - ref_mask).abs() / (base.abs() + 1e-2)).max()) < 2.0 ** -7


@pytest.mark.parametrize("layout", ["oihw", "ohwi"])
@pytest.mark.parametrize("B,H,W,E,Ch,Co", [(3, 16, 16, 128, 32, 64), (2, 24, 20, 128, 64, 128)])
def test_joint_conv_c_code_folding(layout, B, H, W, E, Ch, Co):
    """conv3x3(cat(c broadcast, h)) (model.py:274-279) with the c_code channels folded into a per-sample border-class
    bias: forward vs F.conv2d on the concatenation (rel err <= 5e-3, bf16 h / weights), and the backward pieces
    dc, dW[:, :E] (fp32 reductions, rel err <= 1e-4 against torch on the same bf16 dy) and dW[:, E:], dh (<= 5e-3)."""
    from sg2b200 import ops
    g = torch.Generator().manual_seed(7)
    c = torch.randn(B, E, generator=g).cuda()
    h = _bf(torch.randn(B, Ch, H, W, generator=g).cuda())
    w = (torch.randn(Co, E + Ch, 3, 3, generator=g) / ((E + Ch) * 9) ** 0.5).cuda()
    wb = w.clone()
    wb[:, E:] = _bf(w[:, E:])                       # the kernels see the h part in bf16, the c part in fp32
    cat = torch.cat((c[:, :, None, None].expand(B, E, H, W), h), 1)
    ref = F.conv2d(cat, wb, padding=1)
    if layout == "ohwi":
        wm, so, se, st = w.permute(0, 2, 3, 1).contiguous(), 9 * (E + Ch), 1, E + Ch
    else:
        wm, so, se, st = w.contiguous(), 9 * (E + Ch), 9, 1
    wst = (wm, so, se, st)
    wpk = w[:, E:].permute(0, 2, 3, 1).reshape(Co, 9, Ch).contiguous().bfloat16()
    x = h.permute(0, 2, 3, 1).contiguous().bfloat16()
    bias9 = ops.joint_bias(c, wst, Co)
    stats = torch.zeros(2 * Co, device="cuda", dtype=torch.float64)
    y, _ = ops.conv_fprop(ops.CONV3, x, wpk, Co, stats=stats, bias9=bias9)
    yf = y.float().permute(0, 3, 1, 2)
    assert _rel(yf, ref) < 5e-3
    assert _rel(stats[:Co], yf.sum((0, 2, 3))) < 1e-3          # statistics include the bias
    # backward
    dy = _bf(torch.randn(B, Co, H, W, generator=g).cuda())
    cat_r = cat.clone().requires_grad_(True)
    w_r = wb.clone().requires_grad_(True)
    F.conv2d(cat_r, w_r, padding=1).backward(dy)
    dyn = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    S = ops.joint_tap_sums(dyn)
    dc = torch.zeros(B, E, device="cuda")
    dwm = torch.full_like(wm, float("nan"))
    ops.joint_c_bwd(S, c, wst, dc=dc, dw=dwm)
    dw_c = dwm.permute(0, 3, 1, 2)[:, :E] if layout == "ohwi" else dwm[:, :E]
    assert _rel(dc, cat_r.grad[:, :E].sum((2, 3))) < 1e-4
    assert _rel(dw_c, w_r.grad[:, :E]) < 1e-4
    ops.joint_c_bwd(S, c, wst, dw=dwm, dw_accumulate=True)
    dw_c = dwm.permute(0, 3, 1, 2)[:, :E] if layout == "ohwi" else dwm[:, :E]
    assert _rel(dw_c, 2 * w_r.grad[:, :E]) < 1e-4


def test_head_and_stem_layout_kernels():
    from sg2b200 import ops
    g = torch.Generator().manual_seed(2)
    B, S = 3, 16
    y = torch.randn(B, S, S, 32, generator=g).cuda().bfloat16()
    img = ops.head_tanh_fwd(y, B, S, S)
    assert _rel(img, torch.tanh(y[..., :3].float()).permute(0, 3, 1, 2)) < 1e-5
    dimg = torch.randn(B, 3, S, S, generator=g).cuda()
    dy = ops.head_tanh_bwd(dimg, img, 32)
    ref = (dimg * (1 - img * img)).permute(0, 2, 3, 1)
    assert _rel(dy[..., :3].float(), ref) < 4e-3 and float(dy[..., 3:].abs().max()) == 0.0
    # stem: im2col rows reproduce conv4x4 s2 p1; col2im is its adjoint
    im = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
    col = ops.stem_im2col(im)
    w = torch.randn(8, 3, 4, 4, generator=g).cuda()
    wk = w.permute(0, 2, 3, 1).reshape(8, 48)
    out = col.view(-1, 64)[:, :48].float() @ wk.t()
    ref = F.conv2d(_bf(im), w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 8)
    assert _rel(out, ref) < 1e-5
    dcol = torch.randn(1, 1, B * (S // 2) ** 2, 64, generator=g).cuda().bfloat16()
    dimg2 = ops.stem_col2im(dcol, B, S)
    imr = im.clone().requires_grad_(True)
    cols_ref = F.unfold(imr, 4, padding=1, stride=2)                       # (B, 3*16, L) with (c, kh, kw) order
    cols_ref = cols_ref.view(B, 3, 16, -1).permute(0, 3, 2, 1).reshape(-1, 48)   # -> rows (kh*4+kw)*3+c
    (cols_ref * dcol.view(-1, 64)[:, :48].float()).sum().backward()
    assert _rel(dimg2, imr.grad) < 1e-5
    x = torch.randn(B, 4, 4, 64, generator=g).cuda().bfloat16()
    flat = ops.nhwc_to_nchw_f32(x)
    assert torch.equal(flat, x.float().permute(0, 3, 1, 2).reshape(B, -1))
    back = ops.nchw_f32_to_nhwc(flat, B, 4, 4, 64)
    assert torch.equal(back, x)


def test_small_fp32_ops():
    from sg2b200 import ops
    g = torch.Generator().manual_seed(3)
    M, K1, K2, N, E = 6, 128, 100, 512, 128
    x1, x2 = torch.randn(M, K1, generator=g).cuda(), torch.randn(M, K2, generator=g).cuda()
    w, b = torch.randn(N, K1 + K2, generator=g).cuda() * 0.05, torch.randn(N, generator=g).cuda()
    xr = torch.cat((x1, x2), 1).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.linear(xr, wr, br)
    out = ops.linear_fwd(x1, x2, w, b, False)
    assert _rel(out, ref) < 1e-5
    dy = torch.randn(M, N, generator=g).cuda()
    dxr, dwr, dbr = torch.autograd.grad(ref, (xr, wr, br), dy)
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.linear_bwd_w(dy, x1, x2, dw, db)
    assert _rel(dw, dwr) < 1e-5 and _rel(db, dbr) < 1e-5
    assert _rel(ops.linear_bwd_x(dy, w, K1), dxr[:, :K1]) < 1e-5
    # CA_NET tail
    fc = torch.randn(M, 4 * E, generator=g).cuda().requires_grad_(True)
    eps = torch.randn(M, E, generator=g).cuda()
    gl = fc[:, :2 * E] * torch.sigmoid(fc[:, 2 * E:])
    mu_r, lv_r = gl[:, :E], gl[:, E:]
    c_r = eps * torch.exp(0.5 * lv_r) + mu_r
    mu, lv, c = ops.ca_glu_reparam_fwd(fc.detach(), eps)
    assert _rel(mu, mu_r) < 1e-5 and _rel(lv, lv_r) < 1e-5 and _rel(c, c_r) < 1e-5
    d1, d2, d3 = (torch.randn(M, E, generator=g).cuda() for _ in range(3))
    ref_g, = torch.autograd.grad((mu_r * d1).sum() + (lv_r * d2).sum() + (c_r * d3).sum(), fc)
    assert _rel(ops.ca_glu_reparam_bwd(fc.detach(), eps, d1, d2, d3), ref_g) < 1e-4
    # D logits (conv k4 s4 + bias + sigmoid on a 4x4 map)
    B, C = 5, 64
    x = torch.randn(B, 4, 4, C, generator=g).cuda().bfloat16()
    wl = (torch.randn(1, C, 4, 4, generator=g) * 0.05).cuda().requires_grad_(True)
    bl = torch.randn(1, generator=g).cuda().requires_grad_(True)
    xn = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    pr = torch.sigmoid(F.conv2d(xn, wl, bl, stride=4)).view(-1)
    p = ops.logits_fwd(x, wl.detach(), bl.detach())
    assert _rel(p, pr) < 1e-5
    dp = torch.randn(B, generator=g).cuda()
    gx, gw, gb = torch.autograd.grad(pr, (xn, wl, bl), dp)
    dx = torch.empty_like(x)
    dw, db = torch.zeros_like(wl), torch.zeros_like(bl)
    ops.logits_bwd(dp, p, x, wl.detach(), dx, False, dw, db)
    assert _rel(dx.float().permute(0, 3, 1, 2), gx) < 4e-3 and _rel(dw, gw) < 1e-5 and _rel(db, gb) < 1e-5


def test_losses_and_adam():
    from sg2b200 import ops
    from oracle.stackgan_oracle import class_aware_loss, kl_loss
    g = torch.Generator().manual_seed(4)
    B = 12
    probs = torch.rand(6, B, generator=g).cuda().clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    tg, wt = [1, 1, 0, 1, 0, 0], [1, .5, 1, .5, 1, .5]
    ref = sum(w * F.binary_cross_entropy(probs[i], torch.full((B,), float(t), device="cuda")) for i, (t, w) in enumerate(zip(tg, wt)))
    gref, = torch.autograd.grad(ref, probs)
    loss = torch.zeros(1, device="cuda")
    dp = torch.empty(6, B, device="cuda")
    tg_d, wt_d = torch.tensor(tg, dtype=torch.float32).cuda(), torch.tensor(wt, dtype=torch.float32).cuda()
    ops._call("sg2_gan_bce", 1, probs.data_ptr(), tg_d.data_ptr(), wt_d.data_ptr(), 6, B, loss.data_ptr(),
              dp.data_ptr(), ops._st())
    assert _rel(loss, ref.detach().reshape(1)) < 1e-5 and _rel(dp, gref) < 1e-4
    mu = torch.randn(B, 128, generator=g).cuda().requires_grad_(True)
    lv = (torch.randn(B, 128, generator=g) * 0.3).cuda().requires_grad_(True)
    ref = kl_loss(mu, lv) * 2.0
    g1, g2 = torch.autograd.grad(ref, (mu, lv))
    loss.zero_()
    dmu, dlv = torch.empty_like(mu), torch.empty_like(lv)
    ops._call("sg2_kl_loss", 1, mu.data_ptr(), lv.data_ptr(), mu.numel(), 2.0, loss.data_ptr(), dmu.data_ptr(), dlv.data_ptr(), ops._st())
    assert _rel(loss, ref.detach().reshape(1)) < 1e-5 and _rel(dmu, g1) < 1e-5 and _rel(dlv, g2) < 1e-5
    for labels in ([0, 1, 0, 2, 1, 0, 3, 3, 0, 5, 6, 7], list(range(B))):          # active / no same-class pair
        x = torch.randn(B, 512, generator=g).cuda().requires_grad_(True)
        ref = class_aware_loss(x, labels)
        gx = torch.autograd.grad(ref.sum(), x, allow_unused=True)[0] if ref.requires_grad else None
        loss.zero_()
        dx = torch.empty_like(x)
        ws = torch.empty(2 * B * B, device="cuda")
        lab = torch.tensor(labels, dtype=torch.int32).cuda()
        ops._call("sg2_cal_loss", 3, x.data_ptr(), lab.data_ptr(), B, 512, ws.data_ptr(), loss.data_ptr(), dx.data_ptr(), ops._st())
        assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        if gx is not None and float(ref) > 0:
            assert _rel(dx, gx) < 1e-4
        else:
            assert float(dx.abs().max()) == 0.0
    # Adam + EMA vs torch.optim.Adam over 3 steps
    n = 1000
    p0 = torch.randn(n, generator=g).cuda()
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=2e-4, betas=(0.5, 0.999))
    p, m, v, avg = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.clone()
    avg_ref = p0.clone()
    step, bc = torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(2, device="cuda")
    for _ in range(3):
        gr = torch.randn(n, generator=g).cuda()
        pr.grad = gr.clone()
        opt.step()
        avg_ref.mul_(0.999).add_(pr.detach(), alpha=0.001)
        ops._call("sg2_adam_tick", 1, step.data_ptr(), bc.data_ptr(), 0.5, 0.999, ops._st())
        ops._call("sg2_adam_ema", 1, p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), avg.data_ptr(), n, 2e-4, 0.5, 0.999, 1e-8, bc.data_ptr(), 0.999, None, ops._st())
    assert _rel(p, pr.detach()) < 1e-6 and _rel(avg, avg_ref) < 1e-6


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("P1,C", [(384, 256), (24000, 64)])
def test_bn_groups_equal_separate_calls(act, P1, C):
    """groups = 3 (the batched real / wrong / fake discriminator pass): statistics, outputs, dx of each sub-batch equal
    those of three separate calls up to the fp32 atomic summation order of the statistics (mean / rstd rel <= 1e-6, a
    bf16 output may flip by one ulp: rel <= 1e-3); running statistics see three momentum updates in order; dgamma /
    dbeta are the sums over the groups (rel <= 1e-5)."""
    from sg2b200 import ops
    g = torch.Generator().manual_seed(P1 + C + act)
    x = (torch.randn(3 * P1, C, generator=g) * 1.3 + 0.2).cuda().bfloat16()
    Co = C // 2 if act == 1 else C
    dout = torch.randn(3 * P1, Co, generator=g).cuda().bfloat16()
    gamma = (torch.randn(C, generator=g) * 0.1 + 1).cuda()
    beta = (torch.randn(C, generator=g) * 0.1).cuda()
    rm3, rv3, nbt3 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    st3 = ops.bn_stats32(C, x.device, 3)
    ops.bn_stats(x, st3, 3)
    out3, mean3, rstd3 = ops.bn_act_fwd(x, gamma, beta, act, stats=st3, running=(rm3, rv3, nbt3), groups=3)
    dg3, db3 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx3 = ops.bn_act_bwd(x, dout, mean3, rstd3, gamma, beta, act, dg3, db3, False, groups=3)
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    for k in range(3):
        xs, ds = x[k * P1:(k + 1) * P1].contiguous(), dout[k * P1:(k + 1) * P1].contiguous()
        st = ops.bn_stats32(C, x.device)
        ops.bn_stats(xs, st)
        o, m, r = ops.bn_act_fwd(xs, gamma, beta, act, stats=st, running=(rm, rv, nbt))
        assert _rel(o.float(), out3[k * P1:(k + 1) * P1].float()) < 1e-3
        assert _rel(m, mean3.view(3, C)[k]) < 1e-6 and _rel(r, rstd3.view(3, C)[k]) < 1e-6
        dxk = ops.bn_act_bwd(xs, ds, m, r, gamma, beta, act, dg, db, k > 0)
        assert _rel(dxk.float(), dx3[k * P1:(k + 1) * P1].float()) < 1e-3
    assert _rel(rm3, rm) < 1e-6 and _rel(rv3, rv) < 1e-6 and int(nbt3) == int(nbt) == 3
    assert _rel(dg3, dg) < 1e-5 and _rel(db3, db) < 1e-5


def test_pair_kernels_switched_off():
    """SG2_PAIR=0 (read once per process) hands the pair-shaped layers back to the cluster / plain gather kernels: the
    same parity cases must hold on that path too — it is the fallback if the CTA-pair kernels ever have to be disabled
    again (profiles/r02_pair_deadlock.md). Runs in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SG2_PAIR="0")
    sel = ("test_conv_fprop_dgrad_wgrad and (2-12-16-16-128-256 or 0-20-4-4-256-512 or 1-24-4-4-256-256) "
           "or test_conv_dgrad_epilogue_operand and 0-20-4-4-256-512")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), "-q", "-x", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "4 passed" in r.stdout, r.stdout[-500:]
