import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sg2b200 import config
config.set_precision(os.environ.get("PREC", "fp32"))
from tests.parity_util import bucket_grads, build_trainer_and_oracles, loss_vector, oracle_grads, oracle_step, rel, train_batch
branches, B = int(sys.argv[1]), int(sys.argv[2])
LR = float(os.environ.get("LR", "0"))
cfg, ocfg, netG, netsD, tr, (orc32, orc) = build_trainer_and_oracles(branches, seed=0, n_oracles=2, lr=LR)
from tests.parity_util import f64_state
from oracle.stackgan_oracle import OracleTrainer
orc = OracleTrainer(ocfg, f64_state({k: v.detach() for k, v in orc.g.items()}), [f64_state({k: v.detach() for k, v in d.items()}) for d in orc.ds], device="cuda")
b = train_batch(cfg, B, 11)
losses = tr.step(b["z"], b["emb"], b["real"], b["wrong"], b["labels"], eps=b["eps"]).cpu().tolist()
o = oracle_step(orc, b, keep_grads=True)
print("losses", [f"{abs(a-r)/(abs(r)+1e-12):.1e}" for a, r in zip(losses, loss_vector(o))])
o32 = oracle_step(orc32, b, keep_grads=True)
ours, ref, y = bucket_grads(tr), oracle_grads(o), oracle_grads(o32)
for k in ours:
    print(f"{k:45s} ours {rel(ours[k], ref[k]):.2e}  fp32-oracle {rel(y[k], ref[k]):.2e}")
for net in ["G"] + [f"D{i}" for i in range(branches)]:
    f = lambda d: torch.cat([v.detach().double().flatten() for k, v in d.items() if k.startswith(net + ".")])
    print("FLAT", net, f"ours {rel(f(ours), f(ref)):.2e} fp32-oracle {rel(f(y), f(ref)):.2e}")
