#!/usr/bin/env python
"""BASELINE.json configs[2] without a move generator: G synthetic "games" x T search threads, each thread doing what
ChessPlayer.predict does (pipe.send(planes); pipe.recv()) against ChessModelAPI + the B200 evaluator, then one
PUCT select over a synthetic node.  Reports pipe-inclusive positions/s, mean batch size seen by the evaluator,
and the evaluator's share of the wall time.  (python-chess is not installed; the tree driver is SURVEY 8 f1.)"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=5.0)
    args = ap.parse_args()
    import chess_alpha_zero_b200 as caz
    cfg, weights = bench.network(7)
    model = caz.ChessModel(caz.Config("normal"), precision="bf16", max_batch=max(64, args.games * args.threads))
    model._install(cfg, weights)
    pipes = model.get_pipes(args.games * args.threads)
    planes = bench.synthetic_batch(256, seed=5)
    stop = threading.Event()
    counts = [0] * len(pipes)

    def client(i, pipe):
        rng = np.random.RandomState(i)
        while not stop.is_set():
            x = planes[rng.randint(256)]
            pipe.send(x)
            p, v = pipe.recv()
            idx = rng.permutation(1968)[:30]
            pri = p[idx] / (1e-8 + p[idx].sum())           # prior push (player_chess.py:267-275)
            int(np.argmax(pri * np.sqrt(1.0)))             # stand-in for the host-side select
            counts[i] += 1

    ts = [threading.Thread(target=client, args=(i, p), daemon=True) for i, p in enumerate(pipes)]
    [t.start() for t in ts]
    time.sleep(1.0)
    c0, b0, t0 = sum(counts), model.api.batches, time.perf_counter()
    time.sleep(args.seconds)
    c1, b1, t1 = sum(counts), model.api.batches, time.perf_counter()
    stop.set()
    print(json.dumps({"games": args.games, "search_threads": args.threads, "clients": len(pipes),
                      "positions_per_s_through_pipes": (c1 - c0) / (t1 - t0),
                      "mean_batch": (c1 - c0) / max(b1 - b0, 1), "batches_per_s": (b1 - b0) / (t1 - t0),
                      "note": "one Python process, GIL-bound client threads + pickle over multiprocessing.Pipe "
                              "(the reference's own transport); the evaluator itself does ~90k positions/s at batch 16"}))
    os._exit(0)


if __name__ == "__main__":
    main()
