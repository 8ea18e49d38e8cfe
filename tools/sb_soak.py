#!/usr/bin/env python
"""Soak test of the small-batch kernel: random batch sizes 1..74 back to back (1..148 with a second argument "wide": the
two-slice form behind caz_set_small_batch_limit(h, 148)), every result compared bit for bit with the
first result seen for that (batch size, offset).  A lost flag, a stale tile or a race in the rendezvous would show up as
a mismatch or as CAZ_ERR_KERNEL (bounded spins trap)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import chess_alpha_zero_b200 as caz
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
cfg, weights = bench.network(7)
wide = len(sys.argv) > 2 and sys.argv[2] == "wide"
top = 148 if wide else 74
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=top)
if wide:
    ev.set_small_batch_limit(148)
x = bench.synthetic_batch(256, seed=3, ev=ev)
rng = np.random.default_rng(0)
seen, calls, t0 = {}, 0, time.time()
while time.time() - t0 < seconds:
    b = int(rng.integers(1, top + 1)); off = int(rng.integers(0, 4)) * (256 - top) // 3
    p, v = ev.predict_on_batch(x[off:off + b])
    key = (b, off)
    if key in seen:
        assert np.array_equal(seen[key][0], p) and np.array_equal(seen[key][1], v), f"mismatch at batch {b} offset {off} after {calls} calls"
    else:
        seen[key] = (p.copy(), v.copy())
    calls += 1
# positions are also independent of their batch: row i of any batch equals row i of the 64-batch at the same offset
for (b, off), (p, v) in seen.items():
    if (72, off) in seen:
        n = min(b, 72)
        assert np.array_equal(p[:n], seen[(72, off)][0][:n]), (b, off)
# launches queued back to back (no host synchronisation in between), mixed sizes, device-resident buffers
import torch
d_x = torch.from_numpy(x).cuda(); d_p = torch.empty((top, 1968), device="cuda"); d_v = torch.empty((top,), device="cuda")
t1, queued = time.time(), 0
while time.time() - t1 < seconds / 3:
    for _ in range(50):
        b = int(rng.integers(1, top + 1))
        ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr()); queued += 1
    ev.sync()
    if (b, 0) in seen:
        assert np.array_equal(d_p[:b].cpu().numpy(), seen[(b, 0)][0]), b
print(f"queued back to back: {queued} launches ok")
print(f"soak ok: {calls} calls, {len(seen)} distinct (batch, offset) pairs, {calls / (time.time() - t0):.0f} calls/s")
