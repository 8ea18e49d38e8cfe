#!/usr/bin/env python
"""Training throughput of the `opt` worker's step on the B200: positions/s of
forward (training-mode BN) + losses + backward + Adam at batch 384 per GPU (configs/normal.py:52), 7x256 net.

    python tools/train_bench.py [--precision bf16|fp32] [--batch 384] [--steps 20]
    python -m torch.distributed.run --nproc-per-node N tools/train_bench.py ...    # data parallel: NCCL all-reduce of the gradients

Under torchrun every rank trains on its own shard; after the all-reduce all ranks hold identical weights (checked), and on
two ranks rank 0 also checks the averaged gradient against the fp32 build run on both shards."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=7)
    ap.add_argument("--filters", type=int, default=256)
    ap.add_argument("--verify", action="store_true",
                    help="before timing: the all-reduced gradient of the first step against one GPU running every rank's shard in turn")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce the whole gradient vector after the backward pass has ended")
    args = ap.parse_args()
    world, rank, local = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    import torch
    torch.cuda.set_device(local)
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import encoder, keras_graph
    cfg = keras_graph.build_config(blocks=args.blocks, filters=args.filters)
    w = keras_graph.init_weights(cfg, seed=0)
    ev = caz.Evaluator(cfg, w, device=local, precision="bf16", max_batch=512)
    B = args.batch
    planes = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(B, seed=50 + rank)))
    ev.close()
    rng = np.random.RandomState(rank)
    pt = np.zeros((B, 1968), np.float32)
    for i in range(B):
        idx = rng.permutation(1968)[:30]
        pt[i, idx] = rng.dirichlet([0.5] * 30)
    vt = rng.uniform(-1, 1, size=B).astype(np.float32)
    def shard(r):
        rs = np.random.RandomState(r)
        p_ = np.zeros((B, 1968), np.float32)
        for i in range(B):
            idx = rs.permutation(1968)[:30]
            p_[i, idx] = rs.dirichlet([0.5] * 30)
        return p_, rs.uniform(-1, 1, size=B).astype(np.float32)

    verify = None
    if args.verify and world > 1:
        # data-parallel gradient = mean over ranks of the per-shard gradients (every rank normalises BatchNormalization over
        # its own shard, like Keras' multi-GPU data parallelism): rank 0 recomputes every shard on one GPU and compares
        import torch.distributed as dist
        ev2 = caz.Evaluator(cfg, w, device=local, precision="bf16", max_batch=512)
        tv = caz.Trainer(cfg, w, device=local, precision=args.precision, max_batch=B)
        tv.forward_backward(planes, pt, vt)
        dist.all_reduce(tv._grad_tensor, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        got = (tv._grad_tensor / world).cpu().numpy()
        tv.close()
        if rank == 0:
            ts = caz.Trainer(cfg, w, device=local, precision=args.precision, max_batch=B, distributed=False)
            acc = np.zeros_like(got, dtype=np.float64)
            for r in range(world):
                pl = ev2.encode(encoder.fens_to_records(encoder.synthetic_fens(B, seed=50 + r)))
                p_, v_ = shard(r)
                ts.forward_backward(pl, p_, v_)
                acc += np.concatenate([g.reshape(-1) for g in ts.gradients().values()])
            ts.close()
            want = acc / world
            verify = {"max_abs_diff": float(np.abs(got - want).max()), "max_abs_gradient": float(np.abs(want).max()),
                      "relative": float(np.abs(got - want).max() / np.abs(want).max())}
        ev2.close()
        # the overlapped step (all-reduce of the upper half under the backward pass of the lower half) against the plain one
        ta = caz.Trainer(cfg, w, device=local, precision=args.precision, max_batch=B)
        tb = caz.Trainer(cfg, w, device=local, precision=args.precision, max_batch=B)
        for _ in range(3):
            ta.train_on_batch(planes, pt, vt)
            tb.forward_backward(planes, pt, vt)
            tb.apply()
        fa = np.concatenate([v.reshape(-1) for v in ta.weights().values()])
        fb_ = np.concatenate([v.reshape(-1) for v in tb.weights().values()])
        ta.close()
        tb.close()
        if rank == 0:
            verify["overlapped_vs_plain_weights_after_3_steps_max_abs_diff"] = float(np.abs(fa - fb_).max())
    tr = caz.Trainer(cfg, w, device=local, precision=args.precision, max_batch=B)
    losses = [tr.train_on_batch(planes, pt, vt)["loss"] for _ in range(3)]
    bench.barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fb = 0.0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        if world > 1 and not args.no_overlap:
            out = tr.train_on_batch(planes, pt, vt)
        else:
            out = tr.forward_backward(planes, pt, vt)
            fb += time.perf_counter() - t1
            tr.apply()
        losses.append(out["loss"])
    dt = bench.max_over_ranks(time.perf_counter() - t0, world)
    # all ranks must hold the same weights after the all-reduced updates
    flat = np.concatenate([v.reshape(-1) for v in tr.weights().values()])
    digest = float(np.abs(flat).sum())
    same = bench.max_over_ranks(digest, world) == -bench.max_over_ranks(-digest, world)
    if rank == 0:
        flops = 3 * bench.flops_per_position(args.blocks, args.filters) * B * world * args.steps / dt
        print(json.dumps({"metric": "training positions/s (forward + backward + Adam)", "value": world * B * args.steps / dt, "n_gpus": world,
                          "batch_per_gpu": B, "steps": args.steps, "ms_per_step": 1e3 * dt / args.steps,
                          "forward_backward_ms": 1e3 * fb / args.steps, "precision": args.precision,
                          "net": f"{args.blocks}x{args.filters}", "approx_tflops": flops / 1e12,
                          "loss_first_last": [losses[0], losses[-1]], "ranks_hold_identical_weights": bool(same),
                          "allreduced_gradient_vs_one_gpu_running_all_shards": verify,
                          "collective": ("torch.distributed.all_reduce (NCCL) of the flat fp32 gradient vector" +
                                         ("" if args.no_overlap else ", upper half under the backward pass of the lower half")) if world > 1 else None}))
    tr.close()


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
