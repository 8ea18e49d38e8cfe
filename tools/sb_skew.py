#!/usr/bin/env python
"""When does each CTA of the small-batch kernel finish each layer (globaltimer, ns)?  Board tiles synchronise only
within themselves, so the kernel is as long as its slowest tile: prints per-layer finish times per tile."""
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
planes = bench.synthetic_batch(64, seed=3, ev=ev)
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
t0 = st[16 * 8 + 0]   # globaltimer at CTA 0's start
print("CTAs:", ncta, " kernel start (CTA 0) = 0")
for l in (0, 1, 7, 14):
    j = (g[l, :ncta, 0] - t0) * 1e-3
    print(f"layer {l:2d} finished: min {j.min():7.2f} us  median {np.median(j):7.2f}  max {j.max():7.2f}   (CTA 0: {j[0]:7.2f})")
slices = ncta // ((b + 1) // 2) if b > 16 else ncta // b if ncta % b == 0 else 8
j = (g[14, :ncta, 0] - t0) * 1e-3
print("last layer finish per CTA, in block-id order:")
for i in range(0, ncta, 16):
    print("   ", " ".join(f"{x:6.1f}" for x in j[i:i + 16]))

smid = g[15, :ncta, 0].astype(int)
order = np.argsort(smid)
print("by SM id: (smid, tile, finish us)")
for i in range(0, ncta, 8):
    print("   ", "  ".join(f"{smid[k]:3d}/t{k // slices:02d}/{j[k]:5.1f}" for k in order[i:i + 8]))
