#!/usr/bin/env python
"""BASELINE.json configs[1]: batch sweep 1..4096 on one B200 (device-resident timing, CUDA events on the
evaluator's stream; host-visible call latency with pinned buffers).  Prints one JSON line per batch size."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,48,64,128,256,512,1024,2048,4096")
    ap.add_argument("--blocks", type=int, default=7)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--small-limit", type=int, default=-1)
    args = ap.parse_args()
    import torch
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import _lib
    batches = [int(b) for b in args.batches.split(",")]
    cfg, weights = bench.network(args.blocks)
    ev = caz.Evaluator(cfg, weights, precision="bf16", max_batch=max(batches))
    ev.set_small_batch_limit(args.small_limit)
    stream = torch.cuda.ExternalStream(ev.stream)
    planes = bench.synthetic_batch(max(batches), seed=3, ev=ev)
    d_x = torch.from_numpy(planes).cuda()
    d_p = torch.empty((max(batches), 1968), dtype=torch.float32, device="cuda")
    d_v = torch.empty((max(batches),), dtype=torch.float32, device="cuda")
    for b in batches:
        iters = max(20, min(args.iters, 200000 // max(b, 1)))
        for _ in range(10):
            ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
        ev.sync()
        dev = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
            e1.record(stream)
            e1.synchronize()
            dev.append(e0.elapsed_time(e1) * 1e3)
        # the library's own events, recorded right before / after its launches (no Python between event and launch)
        lib_dev = []
        ev.profile(True)
        ev.profile_read(reset=True)
        for _ in range(iters):
            ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
            st = ev.profile_read(reset=True)
            lib_dev.append(sum(v["ms"] for v in st.values() if v["launches"] > 0) * 1e3)
        ev.profile(False)
        # back-to-back throughput (launches queued ahead)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            ev.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
        e1.record(stream)
        e1.synchronize()
        thr = b * iters / (e0.elapsed_time(e1) * 1e-3)
        px, pp, pv = _lib.PinnedArray((b, 18, 8, 8), np.float32), _lib.PinnedArray((b, 1968), np.float32), \
            _lib.PinnedArray((b, 1), np.float32)
        px.array[:] = planes[:b]
        host = []
        for _ in range(iters):
            t0 = time.perf_counter()
            ev.predict_into(px.array, pp.array, pv.array)
            host.append((time.perf_counter() - t0) * 1e6)
        # the same call from board records (68 B per position; the encoder runs inside the forward pass)
        from chess_alpha_zero_b200 import encoder
        prec = _lib.PinnedArray((b, 68), np.uint8)
        fens = encoder.synthetic_fens(256, seed=3)
        prec.array[:] = encoder.fens_to_records([fens[i % 256] for i in range(b)])
        pv1 = _lib.PinnedArray((b,), np.float32)
        host_rec = []
        for _ in range(iters):
            t0 = time.perf_counter()
            ev.predict_records_into(prec.array, pp.array, pv1.array)
            host_rec.append((time.perf_counter() - t0) * 1e6)
        print(json.dumps({"batch": b, "device_p50_us": float(np.median(lib_dev)), "device_min_us": float(np.min(lib_dev)),
                          "device_incl_python_launch_p50_us": float(np.median(dev)),
                          "host_call_p50_us": float(np.median(host)), "host_call_records_p50_us": float(np.median(host_rec)), "positions_per_s_back_to_back": thr,
                          "tflops": thr * bench.flops_per_position(args.blocks) / 1e12}), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
