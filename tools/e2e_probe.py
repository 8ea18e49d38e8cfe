#!/usr/bin/env python
"""End-to-end (pinned host in, host out) throughput of Evaluator.predict_into at one batch size."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=B)
x = bench.synthetic_batch(B, seed=1, ev=ev)
px, pp, pv = _lib.PinnedArray((B, 18, 8, 8), np.float32), _lib.PinnedArray((B, 1968), np.float32), _lib.PinnedArray((B, 1), np.float32)
px.array[:] = x
for _ in range(10): ev.predict_into(px.array, pp.array, pv.array)
n = 60
t0 = time.perf_counter()
for _ in range(n): ev.predict_into(px.array, pp.array, pv.array)
dt = (time.perf_counter() - t0) / n
print(f"CAZ_CHUNKS={os.environ.get('CAZ_CHUNKS', 'default'):28s} {dt * 1e3:7.3f} ms/call  {B / dt:10.0f} positions/s")
