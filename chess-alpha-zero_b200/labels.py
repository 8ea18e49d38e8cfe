"""The 1968-entry UCI move-label space the policy head indexes, and the rank-flip permutation.

Product-side statement of Config.labels / n_labels / flipped_labels / unflipped_index / flip_policy
(/root/reference/src/chess_zero/config.py:76-123,143-146,175-185).  The network's policy row is indexed
by these labels; for black to move the client permutes it with ``unflipped_index``
(agent/player_chess.py:232-233).
"""
import numpy as np

_FILES = "abcdefgh"
_KNIGHT_JUMPS = ((-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2))
_PROMOTIONS = "qrbn"


def _name(file, rank):
    return f"{_FILES[file]}{rank + 1}"


def _targets(file, rank):
    """Destination squares of one origin in label order: rank line, file line, rising diagonal, falling
    diagonal (signed step -7..7), then the eight knight jumps; off-board and null moves dropped."""
    line = [(t, rank) for t in range(8)] + [(file, t) for t in range(8)]
    diag = [(file + t, rank + t) for t in range(-7, 8)] + [(file + t, rank - t) for t in range(-7, 8)]
    jumps = [(file + a, rank + b) for a, b in _KNIGHT_JUMPS]
    for tf, tr in line + diag + jumps:
        if 0 <= tf < 8 and 0 <= tr < 8 and (tf, tr) != (file, rank):
            yield tf, tr


def create_uci_labels():
    labels = [_name(f, r) + _name(tf, tr) for f in range(8) for r in range(8) for tf, tr in _targets(f, r)]
    for f in range(8):
        for piece in _PROMOTIONS:
            for tf in (f, f - 1, f + 1):          # push, capture towards a, capture towards h
                if 0 <= tf < 8:
                    labels.append(f"{_FILES[f]}2{_FILES[tf]}1{piece}")
                    labels.append(f"{_FILES[f]}7{_FILES[tf]}8{piece}")
    return labels


def flipped_uci_labels(labels):
    table = str.maketrans("12345678", "87654321")
    return [lab.translate(table) for lab in labels]


LABELS = create_uci_labels()
N_LABELS = len(LABELS)
LABEL_INDEX = {lab: i for i, lab in enumerate(LABELS)}
FLIPPED_LABELS = flipped_uci_labels(LABELS)
UNFLIPPED_INDEX = np.fromiter((LABEL_INDEX[x] for x in FLIPPED_LABELS), dtype=np.int64, count=N_LABELS)


def flip_policy(pol):
    """Config.flip_policy (config.py:175-182) as one gather."""
    return np.asarray(pol)[UNFLIPPED_INDEX]


def legal_mask(moves_uci):
    """uint8[1968] with 1 at the label of every move (what a caller passes with masked softmax)."""
    m = np.zeros(N_LABELS, np.uint8)
    for u in moves_uci:
        m[LABEL_INDEX[u]] = 1
    return m
