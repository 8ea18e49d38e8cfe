"""Multi-GPU layout of the path: independent units, no data-path collective.

Self-play games own their tree and pipes (/root/reference/src/chess_zero/worker/self_play.py:113-154); the only
shared state is read-only weights.  So the partition is games -> ranks round-robin, one evaluator process (own
CUDA context, own ChessModelAPI-compatible server, own copy of the 17.4 MB weight set) per device.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def games_for_rank(n_games, rank, world):
    """game_id mod world == rank."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return list(range(rank, n_games, world))


def split_batch(n_positions, rank, world):
    """Contiguous, near-even [begin, end) slice of a batch for pure-throughput replicas."""
    base, extra = divmod(n_positions, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
