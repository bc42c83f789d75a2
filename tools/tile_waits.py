"""Per-role wait-cycle breakdown of the tile conv kernel for one layer: SG2_TILE_DBG=64 python tools/tile_waits.py G.head2 fprop"""
import os as _os; _os.environ["SG2_PROBES"] = "1"   # diagnostics build: python -m sg2b200.build --probes
import ctypes, os, sys
os.environ["SG2_TILE_DBG"] = str(int(os.environ.get("SG2_TILE_DBG", "0")) | 64)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sg2b200 import _lib
for name, op in [a.split(":") for a in sys.argv[1:]]:
    sys.argv = ["one_layer.py", name, op, "2"]
    exec(open(os.path.join(ROOT, "tools", "one_layer.py")).read())
    buf = (ctypes.c_longlong * (148 * 16))()
    _lib.call("sg2_tile_dbg_read", ctypes.cast(buf, ctypes.c_void_p), 148 * 16)
    import numpy as np
    a = np.array(buf[:]).reshape(148, 16).astype(float)
    m = a.mean(0)
    print(f"{name} {op}: units/CTA {m[10]:.1f}")
    print(f"  producer: total {m[0]:.0f} cyc, wait a_empty {m[1]:.0f}, wait b_empty {m[2]:.0f}")
    print(f"  mma     : total {m[4]:.0f} cyc, wait t_empty {m[5]:.0f}, wait a_full {m[6]:.0f}, wait b_full {m[7]:.0f}")
    print(f"  epilogue: total {m[8]:.0f} cyc, wait t_full {m[9]:.0f}")
