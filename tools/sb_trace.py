#!/usr/bin/env python
"""Per-layer timeline of CTA 0 of the small-batch tower kernel (clock64 stamps -> microseconds at 1.965 GHz)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=160)
ev.set_small_batch_limit(int(os.environ.get("SB_LIMIT", "-1")))   # 148: the two-slice form of 73-148 boards
planes = bench.synthetic_batch(160, seed=3, ev=ev)
d_x = torch.from_numpy(planes).cuda(); d_p = torch.empty((160, 1968), device="cuda"); d_v = torch.empty((160,), device="cuda")
for _ in range(5): ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
lib = _lib.load()
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 1, None, 0))
ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
st = np.zeros(17 * 8, np.int64)
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 0, st.ctypes.data_as(ctypes.c_void_p), st.size))
full = np.zeros(240, np.int64)
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 0, full.ctypes.data_as(ctypes.c_void_p), full.size))
print("warps reach the post-tower barrier at (us):", " ".join(f"{(v - st[0]) / 1965.0:7.2f}" for v in full[200:208]),
      "| leave it at", f"{(full[216] - st[0]) / 1965.0:7.2f}")
print("heads pass 1 (us): start, tile ready, feat, staged, policy+stats, 2nd rendezvous, end:", " ".join(f"{(v - st[0]) / 1965.0:7.2f}" for v in list(st[15*8+2:15*8+8]) + [st[16*8+3]]))
print("heads pass 2 (us):", " ".join(f"{(v - st[0]) / 1965.0:7.2f}" for v in list(full[224:230]) + [full[233]]))
st = st.reshape(17, 8).astype(np.float64)
t0 = st[0, 0]
us = (st - t0) / 1965.0
names = ["barrier", "A0 in", "A3 in", "mma iss", "acc rdy", "fenced", "joined"]
print("layer " + " ".join(f"{n:>9s}" for n in names) + "   (us since layer-0 start)")
print("kernel: start, pack done, tower done, tile ready, feat done, head weights staged, policy+stats done, 2nd rendezvous:", " ".join(f"{v:8.2f}" for v in us[15]))
cyc = st[16, 2] - st[15, 0]; ns = st[16, 1] - st[16, 0]
print(f"CTA 0 lifetime: {cyc:.0f} cycles in {ns / 1e3:.2f} us (globaltimer) -> SM clock {cyc / ns * 1e3:.0f} MHz; the table assumes 1965 MHz")
for l in (0, 1, 2, 3, 13, 14):
    print(f"{l:5d} " + " ".join(f"{us[l, i]:9.2f}" for i in range(7)) + f"   last A plane landed {us[l, 7]:6.2f}")

