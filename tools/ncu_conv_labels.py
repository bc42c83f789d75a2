"""Labelled ncu records of every convolution launch of one train step.

    # on the GPU box (the plain run must exit 0 first):
    python tools/ncu_conv_labels.py run gpurun_out/conv_labels.json
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
        --clock-control none -k regex:"tile_conv|tile_wgrad|igemm" --csv --log-file gpurun_out/conv_ncu.csv \
        python tools/ncu_conv_labels.py run gpurun_out/conv_labels.json
    # here:
    python tools/ncu_conv_labels.py merge gpurun_out/conv_labels.json gpurun_out/conv_ncu.csv profiles/r02_conv_kernels_ncu.txt profiles/r02_conv_dram_traffic.json

`run` records the conv calls of one real step (ops.profile_begin), then launches exactly those calls once more, in
order, AFTER a marker kernel: the i-th conv kernel ncu sees after the marker is the i-th label. Labels carry the layer
shape, the executed FLOPs and the algorithmic bytes (operands read once + result written once)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KIND = {0: "conv3x3", 1: "upconv3x3", 2: "conv4x4s2", 3: "gemm"}


def algorithmic_bytes(op, kind, B, H, W, Ci, Co):
    """bf16 activations in/out once, bf16 weights once (fp32 for the wgrad result)."""
    Ho, Wo = (2 * H, 2 * W) if kind == 1 else ((H // 2, W // 2) if kind == 2 else (H, W))
    taps = {0: 9, 1: 9, 2: 16, 3: 1}[kind]
    x, y, w = B * H * W * Ci * 2, B * Ho * Wo * Co * 2, taps * Ci * Co * 2
    if op == "wgrad":
        return x + y + 2 * w          # fp32 gradient out
    return x + y + w


def run(labels_path):
    import torch
    from sg2b200 import config, ops, trainer, utils
    B = int(os.environ.get("B", "24"))
    cfg = config.cfg
    torch.manual_seed(0)
    netG, netsD = utils.build_networks(cfg, "cuda")
    tr = trainer.FusedTrainer(netG, netsD, cfg)
    b = utils.synthetic_batch(cfg, B, seed=1, device="cuda")
    for _ in range(2):
        tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"])
    torch.cuda.synchronize()
    ops.profile_begin()
    tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"])
    torch.cuda.synchronize()
    calls, keep = ops.profile_end()
    labels = []
    for name, a, fl in calls:
        op = name.replace("sg2_conv_", "")
        if op == "wgrad":
            kind, Bn, H, W, Ci, Co, sk = a[0], a[4], a[5], a[6], a[7], a[8], a[9]
        else:
            kind, Bn, H, W, Ci, Co, sk = a[0], a[5], a[6], a[7], a[8], a[9], a[10]
        labels.append({"op": op, "kind": KIND[kind], "B": Bn, "H": H, "W": W, "Cin": Ci, "Cout": Co, "splitk": sk,
                       "flops": fl, "alg_bytes": algorithmic_bytes(op, kind, Bn, H, W, Ci, Co)})
    json.dump(labels, open(labels_path, "w"))
    # marker: a kernel name ncu's regex cannot match but that we can find by launch order is not needed — the replay below
    # is the LAST group of conv kernels of the process; merge() takes the last len(labels) matching launches
    for name, a, _ in calls:
        ops.replay_call(name, a)
    torch.cuda.synchronize()
    print("replayed", len(calls), "conv launches")


def merge(labels_path, csv_path, txt_out, json_out):
    labels = json.load(open(labels_path))
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    per = {}
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        d = per.setdefault(i, {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("sg2::", ""),
                               "grid": row.get("Grid Size")})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)      # us
        if m.startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[m] = v
    recs = [per[i] for i in sorted(per)][-len(labels):]
    assert len(recs) == len(labels), (len(recs), len(labels))
    rows = []
    for lab, r in zip(labels, recs):
        us = r["gpu__time_duration.sum"]
        dram = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
        rows.append({**lab, "kernel": r["name"], "grid": r["grid"], "us": us, "tflops": lab["flops"] / us / 1e6,
                     "dram_bytes": dram, "dram_over_alg": dram / lab["alg_bytes"],
                     "tensor_pct": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "warps_pct": r.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "regs": r.get("launch__registers_per_thread")})
    tot_us = sum(r["us"] for r in rows)
    with open(txt_out, "w") as f:
        f.write(f"# every conv launch of one train step (batch 24, 3 stages), ncu --clock-control none, cold cache, serialised\n")
        f.write(f"# {len(rows)} launches, {tot_us / 1000:.3f} ms, {sum(r['flops'] for r in rows) / tot_us / 1e6:.1f} TFLOP/s executed; "
                f"DRAM traffic {sum(r['dram_bytes'] for r in rows) / 1e9:.2f} GB vs algorithmic {sum(r['alg_bytes'] for r in rows) / 1e9:.2f} GB\n")
        f.write("#     us   TF/s tensor% warps% regs  DRAM MB  xAlg  op    kind       B    HxW       Cin->Cout  splitk kernel grid\n")
        for r in sorted(rows, key=lambda r: -r["us"]):
            f.write(f"{r['us']:8.1f} {r['tflops']:6.0f} {r['tensor_pct'] or 0:6.1f} {r['warps_pct'] or 0:6.1f} {int(r['regs'] or 0):4d} "
                    f"{r['dram_bytes'] / 1e6:8.1f} {r['dram_over_alg']:5.2f}  {r['op']:5s} {r['kind']:10s} {r['B']:<4d} {r['H']}x{r['W']:<8d} "
                    f"{r['Cin']:5d}->{r['Cout']:<5d} {r['splitk']:<3d} {r['kernel']} {r['grid']}\n")
    fam = {}
    for r in rows:
        k = r["kernel"]
        e = fam.setdefault(k, {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "alg_bytes": 0.0, "flops": 0.0})
        e["launches"] += 1
        for key in ("us", "dram_bytes", "alg_bytes", "flops"):
            e[key] += r[key]
    top = max(fam.items(), key=lambda kv: kv[1]["us"])
    json.dump({"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one capture of every conv launch of one step "
                         "(tools/ncu_conv_labels.py)",
               "launches": len(rows), "dram_bytes_per_launch": sum(r["dram_bytes"] for r in rows) / len(rows),
               "algorithmic_bytes_per_launch": sum(r["alg_bytes"] for r in rows) / len(rows),
               "dominant_kernel": top[0],
               "dominant_kernel_dram_bytes_per_launch": top[1]["dram_bytes"] / top[1]["launches"],
               "dominant_kernel_algorithmic_bytes_per_launch": top[1]["alg_bytes"] / top[1]["launches"],
               "by_kernel": {k: {**v, "tflops": v["flops"] / v["us"] / 1e6} for k, v in fam.items()}}, open(json_out, "w"), indent=1)
    print(open(txt_out).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        merge(*sys.argv[2:6])
