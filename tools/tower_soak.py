#!/usr/bin/env python
"""Soak test of the one-launch tower (conv_tower.cuh): random batch sizes 75..700 back to back through the device-resident
entry point, mixed with small-batch launches, every result compared bit for bit with the per-layer chain's answer for that
batch.  A missed act_ready arrival, a stale scratch tile or a ring bug shows up as a mismatch or as CAZ_ERR_KERNEL.
    python tools/tower_soak.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import chess_alpha_zero_b200 as caz  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=704)
x = bench.synthetic_batch(704, seed=3, ev=ev)
sizes = [75, 76, 100, 148, 149, 255, 256, 295, 296, 297, 300, 511, 592, 593, 700]   # (up to 74 boards the persistent small-batch kernel runs)
ev.set_fused_tower(0)
want = {b: ev.predict_on_batch(x[:b]) for b in sizes}
ev.set_fused_tower(-1)
import torch  # noqa: E402
d_x = torch.from_numpy(x).cuda()
d_p = torch.empty((704, 1968), device="cuda")
d_v = torch.empty((704,), device="cuda")
torch.cuda.synchronize()
rng = np.random.default_rng(0)
t0, launches, checks = time.time(), 0, 0
while time.time() - t0 < seconds:
    for _ in range(40):       # queued back to back, no host synchronisation in between; small batches interleaved
        b = int(rng.choice(sizes)) if rng.random() < 0.8 else int(rng.integers(1, 73))
        ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
        launches += 1
    b = int(rng.choice(sizes))
    ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
    ev.sync()
    launches += 1
    assert np.array_equal(d_p[:b].cpu().numpy(), want[b][0]) and np.array_equal(d_v[:b].cpu().numpy(), want[b][1].reshape(-1)), \
        f"mismatch at batch {b} after {launches} launches"
    checks += 1
print(f"tower soak ok: {launches} launches, {checks} checked bit for bit against the per-layer chain, {launches / (time.time() - t0):.0f} launches/s")
