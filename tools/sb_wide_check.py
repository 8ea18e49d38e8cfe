"""The persistent small-batch kernel beyond 72 boards (two-board tiles x two 128-channel slices, up to 148 boards)
against the one-launch tower that serves those batches by default: same results within the bf16 build's tolerance
(different accumulation shapes), device time per call.

    python tools/sb_wide_check.py [--batches 73,100,128,148] [--iters 300]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="73,100,128,148")
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import encoder, keras_graph
    cfg = keras_graph.build_config(blocks=7, filters=256)
    w = keras_graph.init_weights(cfg, seed=0)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=512)
    for b in [int(v) for v in args.batches.split(",")]:
        x = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(b, seed=3)))
        out = {}
        for name, limit in (("tower", 72), ("small_batch_kernel", 148)):
            ev.set_small_batch_limit(limit)
            l0 = ev.kernel_launches
            p, v = ev.predict_on_batch(x)
            launches = ev.kernel_launches - l0
            for _ in range(20):
                ev.predict_on_batch(x)
            t0 = time.perf_counter()
            for _ in range(args.iters):
                ev.predict_on_batch(x)
            dt = (time.perf_counter() - t0) / args.iters
            out[name] = (p.copy(), v.copy(), dt, launches)
        ev.set_small_batch_limit(-1)
        (p0, v0, t0_, l0_), (p1, v1, t1_, l1_) = out["tower"], out["small_batch_kernel"]
        print(json.dumps({"batch": b, "tower_us_per_call": round(1e6 * t0_, 1), "small_batch_kernel_us_per_call": round(1e6 * t1_, 1),
                          "launches": [l0_, l1_], "max_abs_dp": float(np.abs(p0 - p1).max()), "max_abs_dv": float(np.abs(v0 - v1).max()),
                          "argmax_agree": float((p0.argmax(1) == p1.argmax(1)).mean())}))
    ev.close()


if __name__ == "__main__":
    main()
