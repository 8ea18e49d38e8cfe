"""Worker for tests/test_multi_rank_cpu.py: launched by torch.distributed.run with 2 ranks, gloo backend."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chess_alpha_zero_b200 import sharding  # noqa: E402


def main():
    world, rank, local = bench.dist_setup(2)
    assert dist.get_backend() == "gloo" and world == 2
    assert (rank, world, local) == sharding.rank_world()
    mine = sharding.games_for_rank(11, rank, world)
    sizes = [None, None]
    dist.all_gather_object(sizes, mine)
    assert sorted(sizes[0] + sizes[1]) == list(range(11)) and not set(sizes[0]) & set(sizes[1])
    b, e = sharding.split_batch(4097, rank, world)
    spans = [None, None]
    dist.all_gather_object(spans, (b, e))
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 4097
    # the benchmark's timing reduction: every rank must see the slowest rank's time
    t = bench.max_over_ranks(1.0 + rank, world)
    assert t == 2.0, t
    bench.barrier(world)
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
