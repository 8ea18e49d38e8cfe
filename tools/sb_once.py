#!/usr/bin/env python
"""A handful of small-batch forward passes (one cooperative launch each) for ncu: `python tools/sb_once.py 16 12`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import chess_alpha_zero_b200 as caz
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=64)
planes = bench.synthetic_batch(64, seed=3, ev=ev)
d_x = torch.from_numpy(planes).cuda(); d_p = torch.empty((64, 1968), device="cuda"); d_v = torch.empty((64,), device="cuda")
for _ in range(n):
    ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
ev.sync()
print("ok", float(d_p[:b].sum()))
