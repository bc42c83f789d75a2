"""Compare two layer_bench logs: python tools/cmp_layers.py old.log new.log [ops...]"""
import sys
def load(f):
    d = {}
    for l in open(f):
        p = l.split()
        if len(p) == 7 and p[1] in ('fprop', 'dgrad', 'wgrad', 'bn_fwd', 'bn_bwd'):
            d[(p[0], p[1])] = (float(p[2]), float(p[3]), int(p[4]))
    return d
a, b = load(sys.argv[1]), load(sys.argv[2])
ops = sys.argv[3:] or ['fprop', 'dgrad', 'wgrad']
tot_a = tot_b = 0
for k in a:
    if k[1] in ops and k in b:
        tot_a += a[k][0] * a[k][2]; tot_b += b[k][0] * b[k][2]
        if abs(a[k][0] - b[k][0]) > 0.08 * a[k][0]:
            print(f"{k[0]:12s} {k[1]:6s} {a[k][0]:8.1f} -> {b[k][0]:8.1f} us   {a[k][1]:7.1f} -> {b[k][1]:7.1f} TF/s  x{a[k][2]}")
print(f"total {tot_a/1000:.3f} -> {tot_b/1000:.3f} ms")
