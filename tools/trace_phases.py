"""Read a tools/step_trace.py event list: cut it into steps (marker: the one kl_loss launch per step) and print, for the
middle step, the phase boundaries and the conv launches of the largest discriminator's share of the G step.

    python tools/trace_phases.py gpurun_out/step_trace.json [--list PATTERN]
"""
import collections, json, sys


def steps_of(ev):
    ev = sorted(ev, key=lambda e: e[2])
    kl = [e[2] for e in ev if "kl_loss" in e[0]]
    off = kl[0] - ev[0][2]
    cuts = [t - off - 0.5 for t in kl] + [float("inf")]
    out = []
    for k in range(len(kl)):
        st = [(a, b, c - cuts[k], d) for a, b, c, d in ev if cuts[k] <= c < cuts[k + 1]]
        out.append(st)
    return out


def phases(step):
    end = max(c + d for _, _, c, d in step)
    t_kl = [c for a, _, c, _ in step if "kl_loss" in a][0]
    col2im = [c + d for a, _, c, d in step if "stem_col2im" in a]
    t_join = max(col2im) if col2im else 0.0
    bce = sorted(c for a, _, c, _ in step if "gan_bce" in a)
    return {"g_fwd_end": t_kl, "d_update_bce": bce[:len(bce) // 2], "gstep_bce": bce[len(bce) // 2:], "g_bwd_start": t_join,
            "end": end, "kernel_time": sum(d for *_, d in step), "kernels": len(step)}


if __name__ == "__main__":
    ev = json.load(open(sys.argv[1]))
    st = steps_of(ev)
    s = st[len(st) // 2]
    ph = phases(s)
    print({k: (round(v, 1) if isinstance(v, float) else [round(x, 1) for x in v] if isinstance(v, list) else v) for k, v in ph.items()})
    if "--list" in sys.argv:
        pat = sys.argv[sys.argv.index("--list") + 1]
        lo = float(sys.argv[sys.argv.index("--from") + 1]) if "--from" in sys.argv else 0.0
        for a, b, c, d in s:
            if pat in a and c >= lo:
                print("%4d %-48s %8.1f %7.1f" % (b, a[:48], c, d))
