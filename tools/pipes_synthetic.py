#!/usr/bin/env python
"""BASELINE.json configs[2] without a move generator: G synthetic "games" x T search threads, each thread doing what
ChessPlayer.predict does (pipe.send(planes); pipe.recv()) against ChessModelAPI + the B200 evaluator, then the prior
push of one expanded node.  Reports pipe-inclusive positions/s, the mean batch the evaluator saw and batches/s.
--procs 0 keeps every client thread in the server's process (one GIL); --procs G runs each game in its own spawned
process, the way the reference's self-play pool does (worker/self_play.py:43-48).  --transport pipe is the
reference's carrier, shm the shared-slab one (shm_pipes.py).  (python-chess is not installed; the tree driver is
SURVEY 8 f1.)"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def client_loop(i, pipe, planes, counts, stop):
    rng = np.random.RandomState(i)
    while not stop.is_set():
        pipe.send(planes[rng.randint(len(planes))])
        p, v = pipe.recv()
        idx = rng.permutation(1968)[:30]
        pri = p[idx] / (1e-8 + p[idx].sum())           # prior push (player_chess.py:267-275)
        int(np.argmax(pri))                            # stand-in for the host-side select
        counts[i] += 1


def game_process(g, pipes, counter, stop_flag):
    planes = np.random.RandomState(g).rand(64, 18, 8, 8).astype(np.float32)
    counts = [0] * len(pipes)
    stop = threading.Event()
    ts = [threading.Thread(target=client_loop, args=(i, p, planes, counts, stop), daemon=True)
          for i, p in enumerate(pipes)]
    [t.start() for t in ts]
    while not stop_flag.value:
        time.sleep(0.05)
        counter.value = sum(counts)
    stop.set()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--transport", choices=["pipe", "shm"], default="pipe")
    ap.add_argument("--procs", type=int, default=0, help="0: client threads in this process; else one process per game")
    args = ap.parse_args()
    import bench
    import chess_alpha_zero_b200 as caz
    cfg, weights = bench.network(7)
    n = args.games * args.threads
    model = caz.ChessModel(caz.Config("normal"), precision="bf16", max_batch=max(64, n), transport=args.transport)
    model._install(cfg, weights)
    pipes = model.get_pipes(n)
    if args.procs:
        ctx = mp.get_context("spawn")
        stop_flag = ctx.Value("i", 0)
        counters = [ctx.Value("q", 0) for _ in range(args.games)]
        procs = [ctx.Process(target=game_process, args=(g, pipes[g * args.threads:(g + 1) * args.threads],
                                                        counters[g], stop_flag), daemon=True)
                 for g in range(args.games)]
        [p.start() for p in procs]
        total = lambda: sum(c.value for c in counters)  # noqa: E731
        while total() == 0:
            time.sleep(0.1)
    else:
        planes = bench.synthetic_batch(256, seed=5, ev=model.model)
        stop = threading.Event()
        counts = [0] * n
        ts = [threading.Thread(target=client_loop, args=(i, p, planes, counts, stop), daemon=True)
              for i, p in enumerate(pipes)]
        [t.start() for t in ts]
        total = lambda: sum(counts)  # noqa: E731
    time.sleep(1.0)
    c0, b0, t0 = total(), model.api.batches, time.perf_counter()
    time.sleep(args.seconds)
    c1, b1, t1 = total(), model.api.batches, time.perf_counter()
    if args.procs:
        stop_flag.value = 1
    else:
        stop.set()
    print(json.dumps({"games": args.games, "search_threads": args.threads, "clients": n,
                      "transport": args.transport, "client_processes": args.procs and args.games,
                      "positions_per_s_through_pipes": (c1 - c0) / (t1 - t0),
                      "mean_batch": (c1 - c0) / max(b1 - b0, 1), "batches_per_s": (b1 - b0) / (t1 - t0)}))
    sys.stdout.flush()
    if args.procs:
        [p.join(timeout=2.0) for p in procs]
    os._exit(0)


if __name__ == "__main__":
    main()
