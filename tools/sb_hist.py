#!/usr/bin/env python
"""Distribution of the small-batch forward time (library events around the launch): `python tools/sb_hist.py 16 2000`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import chess_alpha_zero_b200 as caz
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=64)
planes = bench.synthetic_batch(64, seed=3, ev=ev)
d_x = torch.from_numpy(planes).cuda(); d_p = torch.empty((64, 1968), device="cuda"); d_v = torch.empty((64,), device="cuda")
for _ in range(50): ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
ev.sync(); ev.profile(True); ev.profile_read(reset=True)
t = []
for _ in range(n):
    ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
    st = ev.profile_read(reset=True)
    t.append(sum(v["ms"] for v in st.values() if v["launches"] > 0) * 1e3)
t = np.asarray(t)
print("batch", b, "percentiles 1/5/25/50/75/95/99:", " ".join(f"{np.percentile(t, q):.1f}" for q in (1, 5, 25, 50, 75, 95, 99)))
h, e = np.histogram(t, bins=np.arange(np.floor(t.min()), np.floor(t.min()) + 24, 1.0))
print(" ".join(f"{int(x)}:{c}" for x, c in zip(e[:-1], h)))
print("first 40:", " ".join(f"{x:.0f}" for x in t[:40]))
