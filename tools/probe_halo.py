"""EXPERIMENT: does a UMMA descriptor that starts at a shifted row of a TMA-loaded halo tile read the right data?
Runs conv3x3 fprop through sg2_probe_halo_fprop for pitch in {10, 16} x base-offset mode in {0, 1} and compares
with the production kernel and torch. Each case runs in a child process with a timeout."""
import os as _os; _os.environ["SG2_PROBES"] = "1"   # diagnostics build: python -m sg2b200.build --probes
import ctypes, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(B, H, W, Ci, Co, pitch, bo):
    import torch
    import torch.nn.functional as F
    from sg2b200 import _lib, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    x = torch.randn(B, Ci, H, W, device=dev).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device=dev) / (9 * Ci) ** 0.5).bfloat16()
    ref = F.conv2d(x.float(), w.float(), padding=1).permute(0, 2, 3, 1)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wpk = w.permute(0, 2, 3, 1).reshape(Co, 9, Ci).contiguous()
    y = torch.full((B, H, W, Co), float("nan"), device=dev, dtype=torch.bfloat16)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("sg2_probe_halo_fprop", xn.data_ptr(), wpk.data_ptr(), y.data_ptr(), B, H, W, Ci, Co, pitch, bo, st)
    torch.cuda.synchronize()
    d = (y.float() - ref)
    rel = (d.norm() / ref.norm()).item()
    out = {"rel": rel, "nan": int(torch.isnan(y.float()).sum())}
    if rel < 2e-2:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        f = lambda: _lib.call("sg2_probe_halo_fprop", xn.data_ptr(), wpk.data_ptr(), y.data_ptr(), B, H, W, Ci, Co, pitch, bo, st)
        g = lambda: ops.conv_fprop(0, xn, wpk, Co)
        for name, fn in (("halo_us", f), ("prod_us", g)):
            fn(); torch.cuda.synchronize(); e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            out[name] = 1000 * e0.elapsed_time(e1) / 20
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(one(*json.loads(sys.argv[1]))))
        sys.exit(0)
    shapes = [(2, 32, 32, 64, 64), (2, 32, 32, 128, 64), (2, 32, 32, 32, 32), (24, 128, 128, 32, 32), (24, 64, 64, 64, 64)]
    for shp in shapes:
        for pitch in (10, 16):
            for bo in (0, 1):
                c = list(shp) + [pitch, bo]
                try:
                    r = subprocess.run([sys.executable, __file__, json.dumps(c)], capture_output=True, text=True, timeout=120)
                    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                    print(c, lines[-1] if lines else f"CRASH rc={r.returncode} {r.stderr[-400:]}", flush=True)
                except subprocess.TimeoutExpired:
                    print(c, "TIMEOUT", flush=True)
