"""Oracle: FEN -> canonical (18, 8, 8) float32 input planes.

Test infrastructure (see oracle/__init__.py).  Restates
  canon_input_planes /root/reference/src/chess_zero/env/chess_env.py:231-238
  maybe_flip_fen     /root/reference/src/chess_zero/env/chess_env.py:251-265
  aux_planes         /root/reference/src/chess_zero/env/chess_env.py:268-289
  to_planes          /root/reference/src/chess_zero/env/chess_env.py:323-332
  alg_to_coord       /root/reference/src/chess_zero/env/chess_env.py:311-314

Semantics reproduced, quirks included:
  * planes 0-11: one-hot per piece letter in the order K Q R B N P k q r b n p,
    indexed [row from the top of the FEN board text][file];
  * black to move: the board text is mirrored top<->bottom and letter case is
    swapped, the castling field is case-swapped (sorted, which only matters for
    membership tests), while the en-passant field and both clocks are copied
    UNCHANGED -- so plane 17 is not mirrored;
  * planes 12-15: constant 0/1 for 'K', 'Q', 'k', 'q' present in the (possibly
    swapped) castling field;
  * plane 16: the raw halfmove-clock integer, broadcast;
  * plane 17: one-hot at (8 - ep_rank, ep_file).

It also defines the compact "board record" the CUDA encoder consumes (the product
has its own FEN parser; this one exists so tests can cross-check it).
"""
import numpy as np

PIECE_ORDER = "KQRBNPkqrbnp"
_PLANE_OF = {ch: i for i, ch in enumerate(PIECE_ORDER)}


def _board_rows(placement):
    """FEN placement text -> 8 strings of 8 chars ('.' for an empty square), top rank first."""
    rows = []
    for row in placement.split("/"):
        cells = []
        for ch in row:
            if ch.isdigit():
                cells.extend("." * int(ch))
            else:
                cells.append(ch)
        rows.append(cells)
    return rows


def canon_input_planes(fen):
    placement, side, castling, ep, halfmove = fen.split(" ")[:5]
    rows = _board_rows(placement)
    if side == "b":
        rows = [[c.swapcase() if c != "." else c for c in row] for row in reversed(rows)]
        castling = castling.swapcase()
    planes = np.zeros((18, 8, 8), dtype=np.float32)
    for r in range(8):
        for f in range(8):
            c = rows[r][f]
            if c != ".":
                planes[_PLANE_OF[c], r, f] = 1.0
    for i, flag in enumerate("KQkq"):
        if flag in castling:
            planes[12 + i] = 1.0
    planes[16] = np.float32(int(halfmove))
    if ep != "-":
        planes[17, 8 - int(ep[1]), ord(ep[0]) - ord("a")] = 1.0
    return planes


# ---------------------------------------------------------------------------
# Compact board record (what the CUDA encoder reads): 68 bytes per position.
#   bytes 0..63 : square codes in FEN text order (row from the top, file a..h) of the
#                 position AS WRITTEN (not yet canonicalised):
#                 0 = empty, 1..12 = index into "KQRBNPkqrbnp" + 1
#   byte 64     : bit0 = black to move; bits 1..4 = castling flags K, Q, k, q as written
#   byte 65     : en-passant square as row*8+file with row = 8 - rank, or 255 if none
#   bytes 66,67 : halfmove clock, little-endian uint16
# ---------------------------------------------------------------------------
RECORD_BYTES = 68


def fen_to_record(fen):
    placement, side, castling, ep, halfmove = fen.split(" ")[:5]
    rec = np.zeros(RECORD_BYTES, dtype=np.uint8)
    rows = _board_rows(placement)
    for r in range(8):
        for f in range(8):
            c = rows[r][f]
            rec[r * 8 + f] = 0 if c == "." else _PLANE_OF[c] + 1
    flags = 1 if side == "b" else 0
    for i, flag in enumerate("KQkq"):
        if flag in castling:
            flags |= 2 << i
    rec[64] = flags
    rec[65] = 255 if ep == "-" else (8 - int(ep[1])) * 8 + (ord(ep[0]) - ord("a"))
    clock = int(halfmove)
    rec[66] = clock & 0xFF
    rec[67] = (clock >> 8) & 0xFF
    return rec


def record_to_planes(rec):
    """Numpy statement of what the CUDA encoder must do with one record."""
    rec = np.asarray(rec, dtype=np.uint8)
    black = bool(rec[64] & 1)
    planes = np.zeros((18, 8, 8), dtype=np.float32)
    for r in range(8):
        for f in range(8):
            code = int(rec[r * 8 + f])
            if code == 0:
                continue
            p = code - 1
            if black:
                p = (p + 6) % 12          # swap case
                planes[p, 7 - r, f] = 1.0  # mirror rows
            else:
                planes[p, r, f] = 1.0
    k, q, kk, qq = [(rec[64] >> (1 + i)) & 1 for i in range(4)]
    cast = (kk, qq, k, q) if black else (k, q, kk, qq)
    for i in range(4):
        planes[12 + i] = np.float32(cast[i])
    planes[16] = np.float32(int(rec[66]) | (int(rec[67]) << 8))
    if rec[65] != 255:
        planes[17, rec[65] // 8, rec[65] % 8] = 1.0
    return planes


def synthetic_fens(n, seed=0):
    """Seeded legal-looking FENs: 4-32 pieces on distinct squares, random castling
    subset, halfmove clock U{0..99}, ~10 % with an en-passant square, 50 % black to move."""
    rng = np.random.RandomState(seed)
    fens = []
    pool = list("QRRBBNNPPPPPPPP") + list("qrrbbnnpppppppp")
    for _ in range(n):
        npieces = int(rng.randint(4, 33))
        others = [pool[i] for i in rng.permutation(len(pool))[: npieces - 2]]
        chars = ["K", "k"] + others
        squares = rng.permutation(64)[: len(chars)]
        board = ["."] * 64
        for sq, c in zip(squares, chars):
            board[sq] = c
        rows = []
        for r in range(8):
            row, run = "", 0
            for f in range(8):
                c = board[r * 8 + f]
                if c == ".":
                    run += 1
                else:
                    if run:
                        row += str(run)
                        run = 0
                    row += c
            if run:
                row += str(run)
            rows.append(row)
        side = "b" if rng.rand() < 0.5 else "w"
        castling = "".join(c for c in "KQkq" if rng.rand() < 0.5) or "-"
        ep = "-"
        if rng.rand() < 0.1:
            ep = "abcdefgh"[rng.randint(8)] + ("6" if side == "w" else "3")
        clock = int(rng.randint(0, 100))
        fens.append(f"{'/'.join(rows)} {side} {castling} {ep} {clock} {int(rng.randint(1, 120))}")
    return fens


KNOWN_ANSWER_FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1",
    "rnbqkbnr/ppp1pppp/8/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 2",
    "r3k2r/8/8/8/8/8/8/R3K2R b Kq - 37 60",
    "8/5k2/8/8/8/8/2K5/8 w - - 99 120",
]
