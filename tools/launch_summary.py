"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST complete step.

    python tools/launch_summary.py launches.csv [nsteps] [step_rows_out.csv]
"""
import csv, collections, re, sys
path, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        rows.append((int(row["ID"]), re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", ""), v, row.get("Grid Size")))
# step boundary: ca_glu_reparam_fwd runs once per step, a handful of launches (arena fills, CA_NET fc) after its start
marks = [i for i, r in enumerate(rows) if "ca_glu_reparam_fwd_kernel" in r[1]]
if len(marks) >= 2:   # the last COMPLETE step: between the last two markers
    it = rows[marks[-2] - 6:marks[-1] - 6]
else:
    per = len(rows) // nsteps
    it = rows[-per:]
if len(sys.argv) > 3:          # dump the raw CSV rows of that step (header + rows) for profiles/
    ids = {str(r[0]) for r in it}
    with open(sys.argv[3], "w") as f:
        f.write(lines[0])
        for l in lines[1:]:
            if l.split(",", 1)[0].strip('"') in ids:
                f.write(l)
tot = sum(r[2] for r in it)
print(f"{len(rows)} launches total; last step: {len(it)} launches, {tot/1000:.3f} ms of kernel time (cold-cache, serialised)")
agg = collections.defaultdict(lambda: [0, 0.0])
for _, k, v, g in it:
    agg[k][0] += 1
    agg[k][1] += v
print(f"{'us':>10s} {'share':>6s} {'n':>5s}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{t:10.1f} {100*t/tot:5.1f}% {n:5d}  {k[:100]}")
fam = collections.defaultdict(float)
for _, k, v, g in it:
    f = "igemm (tcgen05 conv fprop/dgrad/wgrad)" if ("igemm" in k or "tile_conv" in k or "tile_wgrad" in k) else ("BN/act/elementwise (sg2)" if "sg2::" in k else "torch (fill/add/copy/rng)")
    fam[f] += v
for f, t in sorted(fam.items(), key=lambda x: -x[1]):
    print(f"family {f:45s} {t/1000:8.3f} ms {100*t/tot:5.1f}%")
if "--top" in sys.argv:
    for r in sorted(it, key=lambda r: -r[2])[:40]:
        print(f"{r[2]:9.1f} us id={r[0]} grid={r[3]} {r[1][:90]}")
