"""Generate tests/golden/* by EXECUTING the reference's own code (build container only).

Test infrastructure (see oracle/__init__.py).  Run as
    python oracle/pin_against_reference.py
in a container that has /root/reference.  The GPU box does not have it; the fixtures this
script writes are committed so nothing at test time reads /root/reference.

What runs as-is from /root/reference/src:
  chess_zero.config                (labels, flipped labels, unflipped_index)      config.py:76-185
  chess_zero.env.chess_env         (canon_input_planes, check_current_planes)     chess_env.py:192-348
  chess_zero.agent.player_chess    (ChessPlayer.select_action_q_and_u)            player_chess.py:251-299
  chess_zero.agent.api_chess       (ChessModelAPI; used by tests, nothing to record)
python-chess is not installed, so ``chess`` is stubbed: the encoder never touches it and
PUCT select only needs ``Move.from_uci`` (identity here), ``board.legal_moves`` and
``board.fen()``.  keras / tensorflow cannot be stubbed meaningfully: the network forward is
NOT pinned by this script (parity unpinned, oracle/__init__.py).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    sys.path.insert(0, "/root/reference/src")
    chess = types.ModuleType("chess")
    chess.pgn = types.ModuleType("chess.pgn")
    chess.WHITE, chess.BLACK = True, False

    class Move:
        @staticmethod
        def from_uci(s):
            return s
    chess.Move = Move
    chess.Board = object
    sys.modules["chess"] = chess
    sys.modules["chess.pgn"] = chess.pgn
    from chess_zero.config import Config
    from chess_zero.env import chess_env
    from chess_zero.agent import player_chess
    return Config, chess_env, player_chess


def main():
    from oracle.encoder import KNOWN_ANSWER_FENS, synthetic_fens
    Config, chess_env, player_chess = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- a10: label space -------------------------------------------------------------------
    labels = list(Config.labels)
    unflipped = np.asarray(Config.unflipped_index, dtype="<i4")
    with open(os.path.join(GOLDEN, "labels.json"), "w") as f:
        json.dump({"labels": labels, "flipped_labels": list(Config.flipped_labels),
                   "unflipped_index": unflipped.tolist(),
                   "sha256_labels": hashlib.sha256(",".join(labels).encode()).hexdigest(),
                   "sha256_unflipped_index": hashlib.sha256(unflipped.tobytes()).hexdigest()}, f)
    rng = np.random.RandomState(7)
    pol = rng.rand(4, Config.n_labels).astype(np.float32)
    flipped = np.stack([Config.flip_policy(p) for p in pol])
    np.savez_compressed(os.path.join(GOLDEN, "flip_policy.npz"), policy=pol, flipped=flipped)

    # ---- a1: encoder ------------------------------------------------------------------------
    fens = KNOWN_ANSWER_FENS + synthetic_fens(256, seed=1234)
    planes = np.stack([chess_env.canon_input_planes(f) for f in fens])
    assert planes.dtype == np.float32 and planes.shape == (len(fens), 18, 8, 8)
    for f, p in zip(fens, planes):
        assert chess_env.check_current_planes(f, p)          # the reference's own round-trip checker
    as_u8 = planes.astype(np.uint8)
    assert np.array_equal(as_u8.astype(np.float32), planes)  # 0/1 and clocks <= 255: lossless
    np.savez_compressed(os.path.join(GOLDEN, "encoder.npz"), planes_u8=as_u8)
    with open(os.path.join(GOLDEN, "encoder_fens.json"), "w") as f:
        json.dump({"fens": fens,
                   "sha256_16": [hashlib.sha256(p.tobytes()).hexdigest()[:16] for p in planes]}, f)

    # ---- a9: PUCT select --------------------------------------------------------------------
    cfg = Config("normal")
    player = player_chess.ChessPlayer(cfg, pipes=[])
    pc = cfg.play

    class Board:
        def __init__(self, moves, key):
            self.legal_moves, self._key = moves, key

        def fen(self):
            return f"{self._key} w - - 0 1"

    class Env:
        def __init__(self, moves, key):
            self.board = Board(moves, key)

    cases = []
    rng = np.random.RandomState(99)
    for case in range(240):
        # typical chess nodes have 20-40 legal moves; every 10th case probes a warp-boundary / maximum size
        n_edges = int(rng.randint(1, 61)) if case % 10 else int(rng.choice([1, 2, 31, 32, 33, 64, 65, 218]))
        idx = rng.permutation(Config.n_labels)[:n_edges]
        moves = [Config.labels[i] for i in idx]
        env = Env(moves, f"node{case}")
        state = player_chess.state_key(env)
        stats = player.tree[state]
        is_root = bool(case % 3 == 0)
        mode = case % 4
        policy_row = rng.dirichlet([0.3] * Config.n_labels).astype(np.float32)
        # float64 view of the float32 network output = the numpy-1.x promotion the reference relied on
        stats.p = policy_row.astype(np.float64)
        player.select_action_q_and_u(env, False)   # first call pushes priors (player_chess.py:267-275)
        pushed = np.asarray([stats.a[m].p for m in moves], dtype=np.float64)
        n = np.zeros(n_edges, dtype=np.int64)
        w = np.zeros(n_edges, dtype=np.float64)
        if mode >= 1:
            n = rng.randint(0, 60, size=n_edges)
            w = rng.uniform(-1, 1, size=n_edges) * n
        if mode == 2 and n_edges >= 3:                      # deliberate exact ties: first index must win
            for m in moves:
                stats.a[m].p = float(1.0 / n_edges)
            n[:] = 4
            w[:] = 2.0
        q = np.zeros(n_edges, dtype=np.float64)
        for j, m in enumerate(moves):
            a = stats.a[m]
            a.n, a.w = int(n[j]), float(w[j])
            a.q = float(w[j] / n[j]) if n[j] else 0.0
            q[j] = a.q
        stats.sum_n = int(n.sum()) + (3 if mode == 3 else 0)   # in-flight virtual loss on some sibling
        p_now = np.asarray([stats.a[m].p for m in moves], dtype=np.float64)
        seed = 1000 + case
        np.random.seed(seed)
        noise = np.random.dirichlet([pc.dirichlet_alpha] * n_edges) if is_root else None
        np.random.seed(seed)
        best = player.select_action_q_and_u(env, is_root)
        cases.append({"label_idx": idx.astype(np.int32), "policy_legal": policy_row[idx], "pushed": pushed,
                      "n": n.astype(np.int32), "q": q, "p": p_now, "sum_n": stats.sum_n,
                      "is_root": is_root, "noise": noise, "best": moves.index(best)})
    # CSR-style flat arrays: edges of case i are [offsets[i], offsets[i+1])
    offsets = np.cumsum([0] + [len(c["n"]) for c in cases]).astype(np.int32)
    cat = lambda k: np.concatenate([c[k] for c in cases])
    np.savez_compressed(
        os.path.join(GOLDEN, "puct.npz"), c_puct=pc.c_puct, noise_eps=pc.noise_eps, offsets=offsets,
        label_idx=cat("label_idx"), policy_legal=cat("policy_legal"), pushed=cat("pushed"), n=cat("n"),
        q=cat("q"), p=cat("p"), sum_n=np.asarray([c["sum_n"] for c in cases], np.int32),
        is_root=np.asarray([c["is_root"] for c in cases], np.uint8),
        noise=np.concatenate([c["noise"] if c["noise"] is not None else np.zeros(len(c["n"])) for c in cases]),
        best=np.asarray([c["best"] for c in cases], np.int32))
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
