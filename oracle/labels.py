"""Oracle: the 1968-entry UCI move-label space and the black/white flip index.

Test infrastructure (see oracle/__init__.py).  Restates
  create_uci_labels      /root/reference/src/chess_zero/config.py:93-123
  flipped_uci_labels     /root/reference/src/chess_zero/config.py:76-90
  Config.unflipped_index /root/reference/src/chess_zero/config.py:185
  Config.flip_policy     /root/reference/src/chess_zero/config.py:175-182

Label order (reference semantics): for every from-square in file-major order
(a1, a2, ..., a8, b1, ...), targets along the rank (by target file), along the
file (by target rank), along the rising diagonal, along the falling diagonal
(both by signed step -7..7), then the eight knight jumps in the fixed order
(-2,-1) (-1,-2) (-2,1) (1,-2) (2,-1) (-1,2) (2,1) (1,2); finally the 176
promotion strings grouped per file and per piece q, r, b, n.
"""
import numpy as np

FILES = "abcdefgh"
RANKS = "12345678"
_KNIGHT = ((-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2))


def _sq(f, r):
    return FILES[f] + RANKS[r]


def uci_labels():
    out = []
    for f in range(8):
        for r in range(8):
            rays = [(t, r) for t in range(8)]
            rays += [(f, t) for t in range(8)]
            rays += [(f + t, r + t) for t in range(-7, 8)]
            rays += [(f + t, r - t) for t in range(-7, 8)]
            rays += [(f + df, r + dr) for df, dr in _KNIGHT]
            for tf, tr in rays:
                if (tf, tr) == (f, r) or not (0 <= tf < 8 and 0 <= tr < 8):
                    continue
                out.append(_sq(f, r) + _sq(tf, tr))
    for f in range(8):
        for piece in "qrbn":
            # straight push first, then capture towards the a-file, then towards the h-file;
            # within each, the white-side "2->1" (a black pawn promoting) precedes "7->8"
            for tf in (f, f - 1, f + 1):
                if not 0 <= tf < 8:
                    continue
                out.append(FILES[f] + "2" + FILES[tf] + "1" + piece)
                out.append(FILES[f] + "7" + FILES[tf] + "8" + piece)
    return out


def flipped_labels(labels=None):
    labels = uci_labels() if labels is None else labels
    mirror = {d: str(9 - int(d)) for d in "12345678"}
    return ["".join(mirror.get(ch, ch) for ch in lab) for lab in labels]


LABELS = uci_labels()
N_LABELS = len(LABELS)
LABEL_INDEX = {lab: i for i, lab in enumerate(LABELS)}
UNFLIPPED_INDEX = np.asarray([LABEL_INDEX[x] for x in flipped_labels(LABELS)], dtype=np.int32)


def flip_policy(pol):
    """pol[unflipped_index] -- exact gather (config.py:175-182)."""
    return np.asarray(pol)[UNFLIPPED_INDEX]
