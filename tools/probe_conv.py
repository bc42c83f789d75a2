"""GPU probe for the tcgen05 implicit-GEMM convolution entry points (fprop / dgrad / wgrad, all kinds).

Not a pytest: a diagnostic harness meant for `gpurun`. Each case runs in a child process with a timeout so a
protocol bug cannot hang the box. Compares against torch fp32 (TF32 off) on bf16-rounded operands.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONV3, UPCONV, CONV4S2, GEMM = 0, 1, 2, 3
KH_S2 = [[1, 3], [2, 0]]
S_UP = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}


def pack_ref(kind, w):
    import torch
    Co, Ci = w.shape[:2]
    if kind == GEMM:
        w2 = w.reshape(Co, Ci)
        return w2.contiguous(), w2.t().contiguous()
    if kind == CONV3:
        return (w.permute(0, 2, 3, 1).reshape(Co, 9, Ci).contiguous(),
                w.permute(1, 2, 3, 0).reshape(Ci, 9, Co).contiguous())
    if kind == CONV4S2:
        wpk = w.permute(0, 2, 3, 1).reshape(Co, 16, Ci).contiguous()
        wT = torch.empty(4, Ci, 4, Co, device=w.device, dtype=w.dtype)
        for py in range(2):
            for px in range(2):
                for a in range(2):
                    for b in range(2):
                        wT[py * 2 + px, :, a * 2 + b, :] = w[:, :, KH_S2[py][a], KH_S2[px][b]].t()
        return wpk, wT.contiguous()
    if kind == UPCONV:
        wpk = torch.empty(4, Co, 4, Ci, device=w.device, dtype=w.dtype)
        wT = torch.empty(Ci, 16, Co, device=w.device, dtype=w.dtype)
        for py in range(2):
            for px in range(2):
                for a in range(2):
                    for b in range(2):
                        wc = sum(w[:, :, kh, kw] for kh in S_UP[(py, a)] for kw in S_UP[(px, b)])
                        wpk[py * 2 + px, :, a * 2 + b, :] = wc
                        wT[:, (py * 2 + px) * 4 + a * 2 + b, :] = wc.t()
        return wpk.contiguous(), wT.contiguous()
    raise ValueError(kind)


def unpack_wgrad_ref(kind, dwpk, Co, Ci):
    import torch
    if kind == GEMM:
        return dwpk.reshape(Co, Ci, 1, 1)
    if kind == CONV3:
        return dwpk.reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2)
    if kind == CONV4S2:
        return dwpk.reshape(Co, 4, 4, Ci).permute(0, 3, 1, 2)
    g = torch.zeros(Co, Ci, 3, 3, device=dwpk.device)
    d = dwpk.reshape(Co, 2, 2, 2, 2, Ci)  # co, py, px, a, b, ci
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    for kh in S_UP[(py, a)]:
                        for kw in S_UP[(px, b)]:
                            g[:, :, kh, kw] += d[:, py, px, a, b, :]
    return g


def ref_fwd(kind, x, w):
    import torch.nn.functional as F
    if kind == GEMM:
        return F.conv2d(x, w)
    if kind == CONV3:
        return F.conv2d(x, w, padding=1)
    if kind == UPCONV:
        return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    return F.conv2d(x, w, stride=2, padding=1)


def err_report(name, got, ref):
    import torch
    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    rel = (diff.norm() / (ref.norm() + 1e-30)).item()
    mx = diff.max().item()
    nan = int(torch.isnan(got).sum().item())
    out = {"what": name, "rel": rel, "max_abs": mx, "ref_absmax": ref.abs().max().item(), "nan": nan}
    if rel > 2e-2 or nan:
        flat = diff.reshape(-1, diff.shape[-1])
        bad_rows = (flat.max(dim=1).values > 0.05 * ref.abs().max()).nonzero().flatten()
        bad_cols = (flat.max(dim=0).values > 0.05 * ref.abs().max()).nonzero().flatten()
        out["n_bad_rows"] = int(bad_rows.numel())
        out["n_rows"] = int(flat.shape[0])
        out["bad_rows_head"] = bad_rows[:24].tolist()
        out["n_bad_cols"] = int(bad_cols.numel())
        out["bad_cols_head"] = bad_cols[:24].tolist()
        out["got_head"] = got.reshape(-1)[:8].tolist()
        out["ref_head"] = ref.reshape(-1)[:8].tolist()
    return out


def run_case(kind, B, H, W, Ci, Co, splitk, reps):
    import ctypes
    import torch
    from sg2b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234 + kind * 7 + Ci + Co + H)
    k = {CONV3: 3, UPCONV: 3, CONV4S2: 4, GEMM: 1}[kind]
    x = torch.randn(B, Ci, H, W, generator=g).to(dev).bfloat16().float()
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(dev)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y_ref = ref_fwd(kind, x, w.bfloat16().float() if kind != UPCONV else w)
    # UPCONV: the pack pre-sums taps in fp32 then rounds; reference keeps fp32 weights (error stays ~bf16 eps)
    dy = torch.randn(y_ref.shape, generator=g).to(dev).bfloat16().float()
    dx_ref, dw_ref = torch.autograd.grad(y_ref, (x, w), dy)
    wpk, wpkT = pack_ref(kind, w.detach())
    wpk = wpk.bfloat16().contiguous()
    wpkT = wpkT.bfloat16().contiguous()
    # CUDA pack kernel vs the torch restatement above (bit-exact expected)
    wpk_k = torch.empty_like(wpk)
    wpkT_k = torch.empty_like(wpkT)
    wd = w.detach().contiguous()
    _lib.call("sg2_pack_weights", kind, wd.data_ptr(), wpk_k.data_ptr(), wpkT_k.data_ptr(), Co, Ci, Co, Ci, 0,
              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    pack_ok = bool(torch.equal(wpk_k, wpk)) and bool(torch.equal(wpkT_k, wpkT))
    if not pack_ok:
        pack_ok = ((wpk_k.float() - wpk.float()).abs().max().item(), (wpkT_k.float() - wpkT.float()).abs().max().item())
    x_nhwc = x.detach().permute(0, 2, 3, 1).contiguous().bfloat16()
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    Ho, Wo = y_ref.shape[2], y_ref.shape[3]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {"case": dict(kind=kind, B=B, H=H, W=W, Ci=Ci, Co=Co, splitk=splitk), "pack_ok": pack_ok, "checks": []}
    flops = 2.0 * y_ref.numel() * Ci * k * k

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- fprop
    if splitk > 1:
        y = torch.zeros(B, Ho, Wo, Co, device=dev, dtype=torch.float32)
        mode = 1
    else:
        y = torch.full((B, Ho, Wo, Co), float("nan"), device=dev, dtype=torch.bfloat16)
        mode = 0
    f = lambda: _lib.call("sg2_conv_fprop", kind, x_nhwc.data_ptr(), wpk.data_ptr(), y.data_ptr(), mode, B, H, W, Ci, Co, splitk, None, 1, 0, None, st)
    f()
    torch.cuda.synchronize()
    res["checks"].append(err_report("fprop", y.permute(0, 3, 1, 2), y_ref.detach()))
    if reps and splitk == 1:
        ms = timed(f)
        res["fprop_ms"] = ms
        res["fprop_tflops_ref"] = flops / ms / 1e9
    # ---- dgrad
    if splitk > 1:
        dx = torch.zeros(B, H, W, Ci, device=dev, dtype=torch.float32)
    else:
        dx = torch.full((B, H, W, Ci), float("nan"), device=dev, dtype=torch.bfloat16)
    f = lambda: _lib.call("sg2_conv_dgrad", kind, dy_nhwc.data_ptr(), wpkT.data_ptr(), dx.data_ptr(), mode, B, H, W, Ci, Co, splitk, None, 0, st)
    f()
    torch.cuda.synchronize()
    res["checks"].append(err_report("dgrad", dx.permute(0, 3, 1, 2), dx_ref))
    if reps and splitk == 1:
        ms = timed(f)
        res["dgrad_ms"] = ms
        res["dgrad_tflops_ref"] = flops / ms / 1e9
    # ---- wgrad
    jobs = {CONV3: 9, UPCONV: 16, CONV4S2: 16, GEMM: 1}[kind]
    dwpk = torch.zeros(Co, jobs, Ci, device=dev, dtype=torch.float32)
    sk = max(1, splitk)
    wsplit = int(os.environ.get("PROBE_WSPLIT", "0")) or max(1, min(64, (B * Ho * Wo // 64) // 4))
    f = lambda: _lib.call("sg2_conv_wgrad", kind, x_nhwc.data_ptr(), dy_nhwc.data_ptr(), dwpk.data_ptr(), B, H, W, Ci, Co, wsplit, st)
    f()
    torch.cuda.synchronize()
    res["checks"].append(err_report("wgrad", unpack_wgrad_ref(kind, dwpk, Co, Ci).permute(0, 2, 3, 1), dw_ref.permute(0, 2, 3, 1)))
    gk = torch.zeros_like(dw_ref)
    _lib.call("sg2_unpack_wgrad", kind, dwpk.data_ptr(), gk.data_ptr(), Co, Ci, Co, Ci, 0, 0, st)
    torch.cuda.synchronize()
    res["unpack_maxdiff"] = (gk - unpack_wgrad_ref(kind, dwpk, Co, Ci)).abs().max().item()
    if reps:
        ms = timed(f)
        res["wgrad_ms"] = ms
        res["wgrad_tflops_ref"] = flops / ms / 1e9
        res["wsplit"] = wsplit
    return res


CASES = [
    # kind, B, H, W, Ci, Co, splitk, reps
    (GEMM, 1, 1, 256, 64, 32, 1, 0),
    (GEMM, 1, 1, 1024, 128, 128, 1, 0),
    (GEMM, 1, 1, 512, 32, 64, 1, 0),
    (GEMM, 1, 1, 512, 16, 32, 1, 0),
    (CONV3, 2, 16, 16, 64, 64, 1, 0),
    (CONV3, 8, 4, 4, 128, 256, 1, 0),
    (CONV3, 3, 8, 8, 32, 32, 1, 0),
    (CONV3, 2, 32, 32, 160, 64, 1, 0),
    (CONV3, 2, 32, 32, 192, 128, 1, 0),
    (CONV3, 24, 4, 4, 640, 512, 4, 0),
    (UPCONV, 8, 4, 4, 64, 64, 1, 0),
    (UPCONV, 2, 16, 16, 32, 32, 1, 0),
    (UPCONV, 2, 32, 32, 128, 128, 1, 0),
    (CONV4S2, 2, 16, 16, 64, 128, 1, 0),
    (CONV4S2, 8, 8, 8, 128, 256, 1, 0),
    (CONV4S2, 8, 8, 8, 256, 512, 8, 0),
    # production shapes (B=24), timed
    (UPCONV, 24, 32, 32, 128, 128, 1, 5),
    (UPCONV, 24, 4, 4, 1024, 1024, 1, 5),
    (CONV3, 24, 64, 64, 192, 128, 1, 5),
    (CONV3, 24, 128, 128, 32, 64, 1, 5),
    (UPCONV, 24, 128, 128, 32, 32, 1, 5),
    (CONV4S2, 24, 64, 64, 128, 256, 1, 5),
    (CONV4S2, 24, 8, 8, 1024, 2048, 1, 5),
    (CONV3, 24, 4, 4, 2048, 1024, 1, 5),
]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        args = json.loads(sys.argv[2])
        print("RESULT " + json.dumps(run_case(*args)))
        return
    sel = os.environ.get("PROBE_CASES")
    cases = CASES if not sel else [CASES[int(i)] for i in sel.split(",")]
    n_bad = 0
    for c in cases:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one", json.dumps(c)], capture_output=True, text=True,
                               timeout=int(os.environ.get("PROBE_TIMEOUT", "150")))
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                res = json.loads(lines[-1][7:])
                ok = all(ch["rel"] < 2e-2 and not ch["nan"] for ch in res["checks"])
                n_bad += (not ok)
                print(("PASS " if ok else "FAIL ") + json.dumps(res), flush=True)
            else:
                n_bad += 1
                print(f"CRASH case={c} rc={r.returncode}\n--stdout--\n{r.stdout[-1500:]}\n--stderr--\n{r.stderr[-2500:]}", flush=True)
        except subprocess.TimeoutExpired:
            n_bad += 1
            print(f"TIMEOUT case={c} after {time.time()-t0:.0f}s", flush=True)
    print(f"probe_conv: {len(cases) - n_bad}/{len(cases)} cases passed")


if __name__ == "__main__":
    main()
