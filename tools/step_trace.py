"""Timeline of the captured train step as it really runs (all streams, warm caches): torch.profiler / CUPTI kernel
records of a few CUDA-graph replays -> per-kernel totals and a compact event list (name, stream, start, duration).

    python tools/step_trace.py [out.json]      env: B (24), SG2_* switches as usual
"""
import collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from sg2b200 import config, trainer, utils

B = int(os.environ.get("B", "24"))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "step_trace.json")
cfg = config.cfg
dev = torch.device("cuda:0")
torch.manual_seed(0)
netG, netsD = utils.build_networks(cfg, dev)
tr = trainer.FusedTrainer(netG, netsD, cfg)
b = utils.synthetic_batch(cfg, B, seed=1, device=dev)
cap = trainer.CapturedStep(tr, B)
cap.load(b["emb"], b["real"], b["wrong"], b["labels"])
cap.capture()
for _ in range(5):
    cap.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    cap.replay()
e1.record()
torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / 20:.3f} ms (20 replays)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        cap.replay()
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "device_time_total", 0):
        pass
raw = prof.profiler.kineto_results.events() if hasattr(prof, "profiler") else []
for e in raw:
    if e.device_type() == torch.autograd.DeviceType.CUDA:
        evs.append((e.name(), e.device_resource_id(), e.start_ns() / 1000.0, e.duration_ns() / 1000.0))
evs.sort(key=lambda t: t[2])
# keep the LAST replay: split on the largest gaps
if evs:
    t_end = [s + d for _, _, s, d in evs]
    gaps = sorted(((evs[i + 1][2] - max(t_end[:i + 1][-50:]), i) for i in range(len(evs) - 1)), reverse=True)[:2]
    cut = max(i for _, i in gaps) + 1
    last = evs[cut:]
    t0 = last[0][2]
    span = max(s + d for _, _, s, d in last) - t0
    print(f"last replay: {len(last)} kernels, span {span / 1000:.3f} ms, sum of kernel time {sum(d for *_, d in last) / 1000:.3f} ms")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, st, s, d in last:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("sg2::", "")
        agg[k][0] += 1
        agg[k][1] += d
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"{t:9.1f} us {n:4d}  {k[:90]}")
    json.dump([(re.sub(r"\(.*", "", n).replace("void ", "").replace("sg2::", ""), st, round(s - t0, 2), round(d, 2)) for n, st, s, d in last],
              open(out, "w"))
