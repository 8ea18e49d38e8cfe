#!/usr/bin/env python
"""BASELINE.json configs[2] / configs[3]: self-play MCTS, 800 simulations per move, N concurrent games on one B200
(this rank's share of the games when launched under torchrun: games are sharded by rank, no collective).

The search is the host-side batched driver (include/caz_mcts.h, chess-alpha-zero_b200/mcts.py); every leaf is
evaluated on the GPU through caz_eval_records (68-byte board records in, encoder + network on the device).
`--groups G` splits the games into G parts whose rounds alternate, so one part's tree work overlaps the others'
evaluation (G driver threads, one evaluator handle).  Default: 2 groups when the rank has 8 or more host cores, 4 groups
of one worker thread each below that (8 GPUs on a 32-core box: 6.75 M sims/s with 4 x 1, 5.6 M with 2 x 2).

Prints one JSON line: sims/s (completed descents), positions/s through the evaluator, mean batch, evaluator occupancy
(fraction of wall time inside caz_eval_records), moves and games completed, games/hour extrapolated from the moves.
"""
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


def auto_groups(world=1):
    """2 groups with many host cores per rank, 4 with few (the evaluator must never wait for one group's tree work)."""
    return 2 if (os.cpu_count() or 1) // max(1, world) >= 8 else 4


def run(ev, games=256, sims=800, search_threads=16, seconds=20.0, groups=0, host_threads=0, seed=1, rank=0, world=1,
        label=""):
    """Self-play rounds for `seconds` on the evaluator `ev` (max_batch >= games / groups * search_threads)."""
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import _lib, mcts
    groups = groups or auto_groups(world)
    per_group = games // groups
    cap = per_group * search_threads
    cores = os.cpu_count() or 1
    host_threads = host_threads or max(1, cores // max(1, world) // groups)
    mcfg = mcts.MctsConfig.normal(simulation_num_per_move=sims, search_threads=search_threads)
    searches = [caz.SelfPlaySearch(per_group, mcfg, seed=seed * 1000 + rank * 10 + g, n_threads=host_threads)
                for g in range(groups)]
    # CAZ_SELFPLAY_FULL_POLICY=1: read the whole policy rows back (7 872 B per leaf) and push the priors on the host, as in
    # round 1; default: caz_submit_records_priors -- the device pushes them, 8 B per legal move come back
    full_policy = bool(os.environ.get("CAZ_SELFPLAY_FULL_POLICY")) or bool(os.environ.get("CAZ_SELFPLAY_SYNC"))
    pins = [(_lib.PinnedArray((cap, 1968) if full_policy else (1, 1968), np.float32), _lib.PinnedArray((cap,), np.float32))
            for _ in searches]
    csr = [None if full_policy else (_lib.PinnedArray((cap + 1,), np.int32), _lib.PinnedArray((cap * 256,), np.int32),
                                     _lib.PinnedArray((cap * 256,), np.float64)) for _ in searches]

    # warm-up round (first touch, clocks)
    for g, (pp, pv), c in zip(searches, pins, csr):
        if full_policy:
            rec = g.collect()
            ev.predict_records_into(rec, pp.array, pv.array)
            g.apply(pp.array[:len(rec)], pv.array[:len(rec)])
        else:
            rec, off, lab = g.collect_csr(c[0].array, c[1].array)
            ev.collect(ev.submit_records_priors(rec, c[0].array, c[1].array, c[2].array, pv.array))
            g.apply_priors(c[2].array, pv.array)
    base = [g.stats() for g in searches]

    acc = [dict(collect=0.0, eval=0.0, apply=0.0, leaves=0, rounds=0) for _ in searches]
    stop_at = time.perf_counter() + seconds
    ev_lock = threading.Lock()
    busy = [0.0]
    # caz_submit_records / caz_collect: one group's policy read-back (16 MB per 2 048 leaves) runs under another group's
    # forward pass (two staging slots; with four groups the third submission simply queues behind the first's slot).
    # CAZ_SELFPLAY_SYNC=1: one synchronous caz_eval_records at a time.
    pipelined = not os.environ.get("CAZ_SELFPLAY_SYNC")
    inflight, busy_from = [0], [0.0]

    def loop(i):
        g, (pp, pv), a, c = searches[i], pins[i], acc[i], csr[i]
        while time.perf_counter() < stop_at:
            t0 = time.perf_counter()
            rec = g.collect() if full_policy else g.collect_csr(c[0].array, c[1].array)[0]
            t1 = time.perf_counter()
            if pipelined:
                with ev_lock:
                    if inflight[0] == 0:
                        busy_from[0] = time.perf_counter()
                    inflight[0] += 1
                    ticket = ev.submit_records(rec, pp.array, pv.array) if full_policy else \
                        ev.submit_records_priors(rec, c[0].array, c[1].array, c[2].array, pv.array)
                ev.collect(ticket)
                t2 = time.perf_counter()
                with ev_lock:
                    inflight[0] -= 1
                    if inflight[0] == 0:
                        busy[0] += t2 - busy_from[0]
            else:
                with ev_lock:
                    t1b = time.perf_counter()
                    ev.predict_records_into(rec, pp.array, pv.array)
                    t2 = time.perf_counter()
                    busy[0] += t2 - t1b
            if full_policy:
                g.apply(pp.array[:len(rec)], pv.array[:len(rec)])
            else:
                g.apply_priors(c[2].array, pv.array)
            t3 = time.perf_counter()
            a["collect"] += t1 - t0
            a["eval"] += t2 - t1
            a["apply"] += t3 - t2
            a["leaves"] += len(rec)
            a["rounds"] += 1

    t_start = time.perf_counter()
    threads = [threading.Thread(target=loop, args=(i,)) for i in range(len(searches))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    wall = time.perf_counter() - t_start

    tot = {}
    for g, b in zip(searches, base):
        s = g.stats()
        for k in s:
            tot[k] = tot.get(k, 0) + (s[k] - b[k] if k != "tree_nodes_last" else 0)
    leaves = sum(a["leaves"] for a in acc)
    rounds = sum(a["rounds"] for a in acc)
    moves_per_game = 80.0   # a typical self-play game length, only used for the games/hour extrapolation
    out = {
        "workload": f"self-play MCTS, {sims} sims/move, {games} concurrent games, search_threads {search_threads}, {label}",
        "rank": rank, "world": world, "seconds": wall,
        "sims_per_s": tot["descents"] / wall,
        "positions_per_s_through_evaluator": leaves / wall,
        "mean_batch": leaves / max(1, rounds),
        "evaluator_occupancy": busy[0] / wall,
        "host_threads_per_group": host_threads, "groups": groups, "host_cores": cores, "pipelined_evaluator": pipelined,
        "read_back": "full policy rows, priors pushed on the host" if full_policy else
                     "priors of the legal moves, pushed on the device (caz_submit_records_priors)",
        "per_round_ms": {k: 1e3 * sum(a[k] for a in acc) / max(1, rounds) for k in ("collect", "eval", "apply")},
        "moves_played": tot["moves_played"], "games_finished": tot["games_finished"],
        "terminal_hits": tot["terminal_hits"], "withdrawn": tot["withdrawn"],
        "results": {k: tot[k] for k in ("white_wins", "black_wins", "draws", "resignations", "adjudications")},
        "games_per_hour_measured": tot["games_finished"] / wall * 3600,
        "moves_per_hour": tot["moves_played"] / wall * 3600,
        "games_per_hour_at_80_moves": tot["moves_played"] / wall * 3600 / moves_per_game,
    }
    for g in searches:
        g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--search-threads", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--groups", type=int, default=0, help="0 = 2 with >= 8 host cores per rank, else 4")
    ap.add_argument("--host-threads", type=int, default=0, help="worker threads per group (0 = cores / ranks / groups)")
    ap.add_argument("--blocks", type=int, default=7)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()

    from chess_alpha_zero_b200 import sharding
    _, world, rank = sharding.rank_world()
    # configs[3]: --games is this GPU's share; the job's games are numbered globally, game_id mod world == rank
    mine = sharding.games_for_rank(args.games * world, rank, world)
    assert len(mine) == args.games and mine[0] == rank
    import torch
    torch.cuda.set_device(rank)
    import chess_alpha_zero_b200 as caz

    cfg_net, weights = bench.network(args.blocks)
    args.groups = args.groups or auto_groups(world)
    cap = args.games // args.groups * args.search_threads
    ev = caz.Evaluator(cfg_net, weights, device=rank, precision=args.precision, max_batch=cap)
    out = run(ev, args.games, args.sims, args.search_threads, args.seconds, args.groups, args.host_threads, args.seed, rank,
              world, label=f"{args.blocks}x256 net ({args.precision}), random-init weights")
    print(json.dumps(out), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
