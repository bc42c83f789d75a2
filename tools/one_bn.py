"""Run BN+GLU fwd/bwd on one big tensor a few times (for ncu): python tools/one_bn.py [H] [C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sg2b200 import ops
H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
y = torch.randn(24, H, H, C, device=dev).bfloat16()
dout = torch.randn(24, H, H, C // 2, device=dev).bfloat16()
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
st = torch.zeros(2 * C, device=dev, dtype=torch.float64)
ops.bn_stats(y.view(-1, C), st)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    st2 = st.clone()
    out, mean, rstd = ops.bn_act_fwd(y, gamma, beta, ops.ACT_GLU, stats=st2)
    ops.bn_act_bwd(y, dout, mean, rstd, gamma, beta, ops.ACT_GLU, dg, db, False)
torch.cuda.synchronize()
print("ok")
