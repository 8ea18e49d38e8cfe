#!/usr/bin/env python
"""sha256 of policy and value for a few batch sizes (compare two builds / environment switches bit for bit)."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import chess_alpha_zero_b200 as caz
batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "16,96,128,148,150,512").split(",")]
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=max(batches))
planes = bench.synthetic_batch(max(batches), seed=5, ev=ev)
for b in batches:
    p, v = ev.predict_on_batch(planes[:b])
    print(b, hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest()[:16], hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()[:16],
          f"sum {float(p.sum()):.4f}")
