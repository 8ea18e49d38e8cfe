#!/usr/bin/env python
"""Why is the small-batch forward time multi-modal?  For N traced launches: the time between the library's events around
the launch, against what happens inside the kernel (globaltimer stamps of every CTA after every layer)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=64)
planes = bench.synthetic_batch(64, seed=3, ev=ev)
d_x = torch.from_numpy(planes).cuda(); d_p = torch.empty((64, 1968), device="cuda"); d_v = torch.empty((64,), device="cuda")
for _ in range(20): ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
lib = _lib.load()
_lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 1, None, 0))
ev.profile(True); ev.profile_read(reset=True)
n = 1024 + 16 * 256 * 2
rows = []
for it in range(N):
    ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
    st_ev = ev.profile_read(reset=True)
    t_ev = sum(v["ms"] for v in st_ev.values() if v["launches"] > 0) * 1e3
    st = np.zeros(n, np.int64)
    _lib.check(ev._handle, lib.caz_debug_sb_trace(ev._handle, 2, st.ctypes.data_as(ctypes.c_void_p), n))
    g = st[1024:].reshape(16, 256, 2)[:, :, 0].astype(np.float64)
    ncta = int((g[0] > 0).sum())
    t0 = float(st[16 * 8 + 0]); t_end0 = float(st[16 * 8 + 1])
    fin = (g[:15, :ncta] - t0) * 1e-3        # [layer][cta] us since CTA 0's start
    rows.append((t_ev, fin, (t_end0 - t0) * 1e-3))
t_ev = np.array([r[0] for r in rows])
l14 = np.array([r[1][14].max() for r in rows]); l0 = np.array([r[1][0].max() for r in rows]); life0 = np.array([r[2] for r in rows])
print(f"batch {b}: events p50 {np.median(t_ev):.1f}  | last layer done (slowest CTA) p50 {np.median(l14):.1f} | CTA 0 lifetime p50 {np.median(life0):.1f}")
print("corr(events, slowest CTA's layer-14 time) =", round(float(np.corrcoef(t_ev, l14)[0, 1]), 3),
      " corr(events, layer-0 time) =", round(float(np.corrcoef(t_ev, l0)[0, 1]), 3))
print("events - layer14: p5/p50/p95", " ".join(f"{np.percentile(t_ev - l14, q):.1f}" for q in (5, 50, 95)))
order = np.argsort(t_ev)
fast, slow = rows[order[N // 20]], rows[order[-N // 20]]
for name, r in (("fast", fast), ("slow", slow)):
    fin = r[1]
    worst = int(np.argmax(fin[14]))
    print(f"{name}: events {r[0]:.1f} us; slowest CTA {worst} (tile {worst // 16}); its layer finish times:", " ".join(f"{x:.1f}" for x in fin[:, worst]))
    print(f"      per-layer max over CTAs:", " ".join(f"{x:.1f}" for x in fin.max(axis=1)))
    print(f"      per-layer min over CTAs:", " ".join(f"{x:.1f}" for x in fin.min(axis=1)))
    print(f"      last-layer finish per tile (max over its CTAs):", " ".join(f"{fin[14, i:i + 16].max():.1f}" for i in range(0, fin.shape[1], 16)))
