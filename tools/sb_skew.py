#!/usr/bin/env python
"""Arrival skew at the grid barrier of the small-batch kernel: per layer, when each CTA joined (globaltimer) and
when it saw the barrier open."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=64)
planes = bench.synthetic_batch(64, seed=3)
d_x = torch.from_numpy(planes).cuda(); d_p = torch.empty((64, 1968), device="cuda"); d_v = torch.empty((64,), device="cuda")
for _ in range(5): ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
lib = _lib.load()
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 1, None, 0))
ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
n = 1024 + 16 * 256 * 2
st = np.zeros(n, np.int64)
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 0, st.ctypes.data_as(ctypes.c_void_p), n))
g = st[1024:].reshape(16, 256, 2).astype(np.float64)
ncta = int((g[0, :, 0] > 0).sum())
print("CTAs:", ncta)
for l in range(0, 14):
    j = g[l, :ncta, 0]; o = g[l + 1, :ncta, 1]
    print(f"layer {l:2d}: joined min..max spread {1e-3 * (j.max() - j.min()):6.2f} us | last join -> first open {1e-3 * (o.min() - j.max()):6.2f} us"
          f" | open spread {1e-3 * (o.max() - o.min()):6.2f} us | layer period {1e-3 * (g[l + 1, :ncta, 0].max() - j.max()):6.2f} us")
