"""ctypes binding of include/caz_mcts.h: the host-side batched self-play search.

``SelfPlaySearch`` stands where the reference has ``ChessPlayer`` threads + ``self_play_buffer``
(agent/player_chess.py:119-333, worker/self_play.py:113-154): it owns many concurrent games, hands out the leaves of
one round of descents as 68-byte board records and takes the network's policy/value back.  The evaluator is the
caller's business (``Evaluator.eval_records`` on a B200; a stub in CPU tests) -- this module never touches CUDA.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from . import labels as _labels

RECORD_BYTES = 68

SYMBOLS = (
    "caz_mcts_create", "caz_mcts_destroy", "caz_mcts_collect", "caz_mcts_apply", "caz_mcts_collect_csr",
    "caz_mcts_apply_priors", "caz_mcts_get_stats",
    "caz_mcts_set_position", "caz_mcts_game_fen", "caz_mcts_game_result", "caz_mcts_root_edges",
    "caz_mcts_drain_play_data", "caz_mcts_debug_dirichlet", "caz_chess_perft",
    "caz_chess_legal_moves", "caz_chess_play", "caz_chess_record",
)


class MctsConfig(ctypes.Structure):
    """PlayConfig (configs/normal.py:28-43); defaults are the reference's `normal` values."""
    _fields_ = [("simulation_num_per_move", ctypes.c_int32), ("search_threads", ctypes.c_int32),
                ("c_puct", ctypes.c_double), ("noise_eps", ctypes.c_double), ("dirichlet_alpha", ctypes.c_double),
                ("tau_decay_rate", ctypes.c_double), ("virtual_loss", ctypes.c_int32), ("resign_enabled", ctypes.c_int32),
                ("resign_threshold", ctypes.c_double), ("min_resign_turn", ctypes.c_int32),
                ("max_game_length", ctypes.c_int32), ("n_labels", ctypes.c_int32), ("continuous", ctypes.c_int32),
                ("record_play_data", ctypes.c_int32)]

    @classmethod
    def normal(cls, **over):
        c = cls(simulation_num_per_move=800, search_threads=16, c_puct=1.5, noise_eps=0.25, dirichlet_alpha=0.3,
                tau_decay_rate=0.99, virtual_loss=3, resign_enabled=1, resign_threshold=-0.8, min_resign_turn=5,
                max_game_length=1000, n_labels=_labels.N_LABELS, continuous=1, record_play_data=0)
        for k, v in over.items():
            setattr(c, k, v)
        return c


class MctsStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("descents", "leaves_evaluated", "terminal_hits", "withdrawn", "moves_played", "games_finished",
                 "white_wins", "black_wins", "draws", "resignations", "adjudications", "tree_nodes_last")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def library_path():
    return _build.MCTS_LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if _build.mcts_is_stale():
        _build.build_mcts()
    lib = ctypes.CDLL(_build.MCTS_LIB_PATH)
    P = ctypes.POINTER
    lib.caz_mcts_create.argtypes = [P(MctsConfig), ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int32, P(ctypes.c_void_p)]
    lib.caz_mcts_destroy.argtypes = [ctypes.c_void_p]
    lib.caz_mcts_destroy.restype = None
    lib.caz_mcts_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, P(ctypes.c_int32)]
    lib.caz_mcts_apply.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.caz_mcts_collect_csr.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, P(ctypes.c_int32), ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int32]
    lib.caz_mcts_apply_priors.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.caz_mcts_get_stats.argtypes = [ctypes.c_void_p, P(MctsStats)]
    lib.caz_mcts_set_position.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_char_p]
    lib.caz_mcts_game_fen.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32]
    lib.caz_mcts_game_result.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.caz_mcts_root_edges.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 4 + [ctypes.c_int32]
    lib.caz_mcts_drain_play_data.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.caz_mcts_drain_play_data.restype = ctypes.c_int64
    lib.caz_mcts_debug_dirichlet.argtypes = [ctypes.c_uint64, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.caz_chess_perft.argtypes = [ctypes.c_char_p, ctypes.c_int32]
    lib.caz_chess_perft.restype = ctypes.c_uint64
    lib.caz_chess_legal_moves.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32]
    lib.caz_chess_play.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32]
    lib.caz_chess_record.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    _lib = lib
    return lib


def move_label_table():
    """int32[64*64*5]: (from*64+to)*5 + promotion (0 none, 1 n, 2 b, 3 r, 4 q) -> label index or -1 (a1 = 0 .. h8 = 63)."""
    t = np.full(64 * 64 * 5, -1, np.int32)
    for i, lab in enumerate(_labels.LABELS):
        f = (ord(lab[1]) - ord("1")) * 8 + ord(lab[0]) - ord("a")
        to = (ord(lab[3]) - ord("1")) * 8 + ord(lab[2]) - ord("a")
        promo = "?nbrq".index(lab[4]) if len(lab) > 4 else 0
        t[(f * 64 + to) * 5 + promo] = i
    return t


def dirichlet_draws(seed, alpha, n, draws):
    """`draws` samples of the root noise Dirichlet([alpha] * n) from the search's generator (test hook)."""
    out = np.zeros((draws, n), np.float64)
    if load().caz_mcts_debug_dirichlet(seed, alpha, n, draws, out.ctypes.data_as(ctypes.c_void_p)):
        raise ValueError("bad arguments")
    return out


# ---- chess rules (tests, tools) ------------------------------------------------------------------------------
def perft(fen, depth):
    return int(load().caz_chess_perft(fen.encode(), depth))


def legal_moves(fen):
    """Legal moves in python-chess's generation order (what select_action_q_and_u's tie-break sees)."""
    buf = ctypes.create_string_buffer(4096)
    n = load().caz_chess_legal_moves(fen.encode(), buf, len(buf))
    if n < 0:
        raise ValueError("bad FEN")
    return buf.value.decode().split() if n else []


RESULTS = ("*", "1-0", "0-1", "1/2-1/2")


def play(fen, moves=()):
    """Board(fen), push_uci every move -> (board.result(claim_draw=True), board.fen())."""
    out = ctypes.create_string_buffer(256)
    r = load().caz_chess_play(fen.encode(), " ".join(moves).encode(), out, len(out))
    if r < 0:
        raise ValueError("illegal move or bad FEN")
    return RESULTS[r], out.value.decode()


def record(fen):
    rec = np.zeros(RECORD_BYTES, np.uint8)
    if load().caz_chess_record(fen.encode(), rec.ctypes.data_as(ctypes.c_void_p)):
        raise ValueError("bad FEN")
    return rec


class SelfPlaySearch:
    """`n_games` concurrent self-play games searched in lock-step rounds.

        search = SelfPlaySearch(256, MctsConfig.normal(), seed=1, n_threads=16)
        while ...:
            records = search.collect()                     # uint8 [n][68], one per leaf
            policy, value = evaluator.eval_records(records)
            search.apply(policy, value)
    """

    def __init__(self, n_games, config=None, seed=0, n_threads=None):
        self.config = config or MctsConfig.normal()
        self.n_games = int(n_games)
        self.capacity = self.n_games * int(self.config.search_threads)
        self._records = np.zeros((self.capacity, RECORD_BYTES), np.uint8)
        self._labels = move_label_table()
        self._unflip = np.ascontiguousarray(_labels.UNFLIPPED_INDEX, dtype=np.int32)
        handle = ctypes.c_void_p()
        threads = n_threads if n_threads is not None else (os.cpu_count() or 1)
        rc = load().caz_mcts_create(ctypes.byref(self.config), self.n_games, seed,
                                    self._labels.ctypes.data_as(ctypes.c_void_p),
                                    self._unflip.ctypes.data_as(ctypes.c_void_p), threads, ctypes.byref(handle))
        if rc:
            raise ValueError(f"caz_mcts_create failed ({rc})")
        self._h = handle
        self._n_last = 0

    def close(self):
        if self._h:
            load().caz_mcts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def collect(self):
        n = ctypes.c_int32()
        rc = load().caz_mcts_collect(self._h, self._records.ctypes.data_as(ctypes.c_void_p), self.capacity, ctypes.byref(n))
        if rc:
            raise RuntimeError(f"caz_mcts_collect failed ({rc})")
        self._n_last = n.value
        return self._records[:n.value]

    MAX_MOVES = 256   # bound on a position's legal moves (csrc/chess/mcts.hpp)

    def collect_csr(self, offsets=None, labels=None):
        """collect() plus the leaves' legal moves as policy labels: returns (records [n][68], offsets int32 [n + 1], labels
        int32 [offsets[n]]) -- the input of Evaluator.submit_records_priors.  `offsets` (capacity + 1) and `labels`
        (capacity * 256) may be caller-owned (pinned) int32 arrays."""
        if offsets is None:
            offsets = np.zeros(self.capacity + 1, np.int32)
        if labels is None:
            labels = np.zeros(self.capacity * self.MAX_MOVES, np.int32)
        n = ctypes.c_int32()
        rc = load().caz_mcts_collect_csr(self._h, self._records.ctypes.data_as(ctypes.c_void_p), self.capacity, ctypes.byref(n),
                                         offsets.ctypes.data_as(ctypes.c_void_p), labels.ctypes.data_as(ctypes.c_void_p),
                                         labels.shape[0])
        if rc:
            raise RuntimeError(f"caz_mcts_collect_csr failed ({rc})")
        self._n_last = n.value
        return self._records[:n.value], offsets[:n.value + 1], labels[:int(offsets[n.value])] if n.value else labels[:0]

    def apply_priors(self, priors, value):
        """apply() from per-edge priors pushed on the device (float64, the order of the last collect_csr)."""
        priors = np.ascontiguousarray(priors, dtype=np.float64)
        value = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
        if value.shape[0] < self._n_last:
            raise ValueError("value does not match the last collect_csr()")
        rc = load().caz_mcts_apply_priors(self._h, priors.ctypes.data_as(ctypes.c_void_p), value.ctypes.data_as(ctypes.c_void_p),
                                          self._n_last)
        if rc:
            raise RuntimeError(f"caz_mcts_apply_priors failed ({rc})")

    def apply(self, policy, value):
        policy = np.ascontiguousarray(policy, dtype=np.float32)
        value = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
        if policy.shape != (self._n_last, self.config.n_labels) or value.shape[0] != self._n_last:
            raise ValueError("policy/value do not match the last collect()")
        rc = load().caz_mcts_apply(self._h, policy.ctypes.data_as(ctypes.c_void_p), value.ctypes.data_as(ctypes.c_void_p),
                                   self._n_last)
        if rc:
            raise RuntimeError(f"caz_mcts_apply failed ({rc})")

    def stats(self):
        s = MctsStats()
        load().caz_mcts_get_stats(self._h, ctypes.byref(s))
        return s.as_dict()

    def set_position(self, game, fen):
        if load().caz_mcts_set_position(self._h, game, fen.encode()):
            raise ValueError("bad game index or FEN")

    def fen(self, game):
        buf = ctypes.create_string_buffer(256)
        if load().caz_mcts_game_fen(self._h, game, buf, len(buf)) < 0:
            raise ValueError("bad game index")
        return buf.value.decode()

    def result(self, game):
        return RESULTS[load().caz_mcts_game_result(self._h, game)]

    def drain_play_data(self, dense=True):
        """Rows of the games finished since the last call, as SelfPlayWorker.flush_buffer would json.dump them
        (self_play.py:90-100): [[fen, policy, z], ...]; policy is the dense n_labels list (dense=True) or {label: p}."""
        lib = load()
        need = lib.caz_mcts_drain_play_data(self._h, None, 0)
        if need <= 0:
            return []
        buf = np.zeros(need, np.uint8)
        got = lib.caz_mcts_drain_play_data(self._h, buf.ctypes.data_as(ctypes.c_void_p), need)
        raw = buf[:got].tobytes()
        rows, i = [], 0
        while i < len(raw):
            j = raw.index(b"\0", i)
            fen = raw[i:j].decode()
            z = int(np.frombuffer(raw, np.int8, 1, j + 1)[0])
            n = int(np.frombuffer(raw[j + 2:j + 4], np.uint16)[0])
            pairs = np.frombuffer(raw[j + 4:j + 4 + 6 * n], np.dtype([("l", "<u2"), ("p", "<f4")]))
            if dense:
                pol = np.zeros(self.config.n_labels, np.float64)
                pol[pairs["l"]] = pairs["p"]
                policy = pol.tolist()
            else:
                policy = {int(a): float(b) for a, b in zip(pairs["l"], pairs["p"])}
            rows.append([fen, policy, z])
            i = j + 4 + 6 * n
        return rows

    def root_edges(self, game):
        """(label, n, q, p) arrays of the root's edges in legal-move order, or None before the root is expanded."""
        cap = 256
        lab, n = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        q, p = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
        k = load().caz_mcts_root_edges(self._h, game, *(a.ctypes.data_as(ctypes.c_void_p) for a in (lab, n, q, p)), cap)
        if k < 0:
            return None
        return lab[:k], n[:k], q[:k], p[:k]
