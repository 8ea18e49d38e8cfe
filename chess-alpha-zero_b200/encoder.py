"""Host side of the input encoder: FEN text -> 68-byte board record (the device kernel's input).

The reference turns a FEN into (18,8,8) float32 planes in Python
(/root/reference/src/chess_zero/env/chess_env.py:231-348).  Here the host only tokenises the FEN into a
compact record; canonicalisation (rank mirror + colour swap for black, castling swap, the un-mirrored
en-passant plane, the raw halfmove clock) and plane expansion run on the device
(csrc/kernels_simt.cuh: record_plane_value), bit-exact with the reference.

Record layout (include/caz_b200.h): bytes 0..63 square codes in FEN text order (rank 8 first), 0 = empty,
1..12 = index into "KQRBNPkqrbnp" + 1; byte 64 bit0 = black to move, bits 1..4 = castling K,Q,k,q as
written; byte 65 = en-passant square (8 - rank)*8 + file, 255 if none; bytes 66-67 = halfmove clock (LE).
"""
import numpy as np

from ._lib import RECORD_BYTES

PIECES = "KQRBNPkqrbnp"
_CODE = {ch: i + 1 for i, ch in enumerate(PIECES)}
_CASTLE_BIT = {"K": 2, "Q": 4, "k": 8, "q": 16}


def fen_to_record(fen, out=None):
    fields = fen.split(" ")
    placement, side, castling, ep, halfmove = fields[0], fields[1], fields[2], fields[3], fields[4]
    rec = np.zeros(RECORD_BYTES, np.uint8) if out is None else out
    rec[:] = 0
    sq = 0
    for ch in placement:
        if ch == "/":
            continue
        if ch.isdigit():
            sq += int(ch)
        else:
            rec[sq] = _CODE[ch]
            sq += 1
    if sq != 64:
        raise ValueError(f"bad FEN placement: {placement!r}")
    flags = 1 if side == "b" else 0
    for ch in castling:
        flags |= _CASTLE_BIT.get(ch, 0)
    rec[64] = flags
    rec[65] = 255 if ep == "-" else (8 - int(ep[1])) * 8 + (ord(ep[0]) - 97)
    clock = int(halfmove)
    if not 0 <= clock < 65536:
        raise ValueError("halfmove clock out of range")
    rec[66], rec[67] = clock & 255, clock >> 8
    return rec


def fens_to_records(fens):
    recs = np.zeros((len(fens), RECORD_BYTES), np.uint8)
    for i, fen in enumerate(fens):
        fen_to_record(fen, recs[i])
    return recs


def synthetic_fens(n, seed=0):
    """Seeded legal-looking FENs for benchmarks: 4-32 pieces on distinct squares, random castling subset,
    halfmove clock U{0..99}, ~10 % with an en-passant square, 50 % black to move."""
    rng = np.random.RandomState(seed)
    pool = list("QRRBBNNPPPPPPPP") + list("qrrbbnnpppppppp")
    fens = []
    for _ in range(n):
        count = int(rng.randint(4, 33))
        chars = ["K", "k"] + [pool[i] for i in rng.permutation(len(pool))[: count - 2]]
        board = ["."] * 64
        for sq, c in zip(rng.permutation(64)[: len(chars)], chars):
            board[sq] = c
        rows = []
        for r in range(8):
            row = "".join(board[r * 8:r * 8 + 8])
            for run in range(8, 0, -1):
                row = row.replace("." * run, str(run))
            rows.append(row)
        side = "b" if rng.rand() < 0.5 else "w"
        castling = "".join(c for c in "KQkq" if rng.rand() < 0.5) or "-"
        ep = "abcdefgh"[rng.randint(8)] + ("6" if side == "w" else "3") if rng.rand() < 0.1 else "-"
        fens.append(f"{'/'.join(rows)} {side} {castling} {ep} {int(rng.randint(0, 100))} {int(rng.randint(1, 120))}")
    return fens
