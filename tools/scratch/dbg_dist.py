import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.multiprocessing as mp
import tests.test_gpu_dist as T
from tests.parity_util import rel

if __name__ == "__main__":
    ctx = mp.get_context("spawn"); port = T._free_port()
    tmp = tempfile.mkdtemp(); out = os.path.join(tmp, "rank0.pt")
    procs = [ctx.Process(target=T._rank_main, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]; [p.join(600) for p in procs]
    real = torch.load(out)
    dev = torch.device("cuda", 0)
    cfg, tr = T._build(dev)
    from sg2b200 import _lib
    _lib.call("sg2_set_sm_reserve", 12)
    hook = T._Hook(tr); tr.all_reduce = hook; nD = len(tr.bD)
    shards = [T._shard(cfg, r, dev) for r in range(2)]
    snap0 = tr.snapshot()
    def run(b):
        tr.restore(snap0)
        o = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).clone(); torch.cuda.synchronize(); return o
    per = []
    for b in shards:
        hook.mode, hook.rec, hook.sub = "record", {}, {}
        run(b); per.append({k: v for k, v in hook.rec.items() if k[0] < nD})
    print("phase1 keys", sorted(per[0].keys())[:20], len(per[0]))
    avg_d = {k: (per[0][k] + per[1][k]) * 0.5 for k in per[0]}
    per = []
    for b in shards:
        hook.mode, hook.rec, hook.sub = "record", {}, dict(avg_d)
        run(b); per.append({k: v for k, v in hook.rec.items() if k[0] == nD})
    avg_g = {k: (per[0][k] + per[1][k]) * 0.5 for k in per[0]}
    hook.mode, hook.rec, hook.sub = "off", {}, {**avg_d, **avg_g}
    losses = run(shards[0]); emu = tr.snapshot()
    print("losses emu", losses.tolist(), "real", real["losses"].tolist())
    names = ["G", "D0", "D1", "D2"]
    for n, be, br in zip(names, emu["buckets"], real["snap"]["buckets"]):
        for k in be:
            if be[k].is_floating_point():
                print(n, k, f"{rel(be[k], br[k].to(dev)):.3e}")
    for n, le, lr in zip(names, emu["buffers"], real["snap"]["buffers"]):
        print(n, "buffers max", max(rel(a.float(), b.to(dev).float()) for a, b in zip(le, lr) if a.is_floating_point()))
