#!/usr/bin/env python
"""Per-stage device time (CUDA events inside the library) for a few batch sizes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
import chess_alpha_zero_b200 as caz

batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "1,16,64").split(",")]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else -1
cfg, weights = bench.network(7)
ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=max(batches))
ev.set_small_batch_limit(limit)
planes = bench.synthetic_batch(max(batches), seed=3, ev=ev)
d_x = torch.from_numpy(planes).cuda()
d_p = torch.empty((max(batches), 1968), dtype=torch.float32, device="cuda")
d_v = torch.empty((max(batches),), dtype=torch.float32, device="cuda")
for b in batches:
    for _ in range(10):
        ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
    ev.sync()
    ev.profile(True); ev.profile_read(reset=True)
    n = 50
    for _ in range(n):
        ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
    st = ev.profile_read(reset=True); ev.profile(False)
    print(json.dumps({"batch": b, "limit": limit, **{k: round(v["ms"] / n * 1e3, 1) for k, v in st.items()}, "unit": "us"}))
