#!/usr/bin/env python
"""Headline benchmark: positions/sec of the batched policy/value evaluation (7x256 net, batch 4096 per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = one forward pass over one batch of synthetic positions (random-init weights of the named
architecture; there are no trained weights or datasets offline).  One process per GPU (torchrun sets
RANK/LOCAL_RANK/WORLD_SIZE); games/positions are independent, so ranks share nothing on the data path
(weak scaling, no collective); torch.distributed is used only for the barrier and the max-over-ranks time.

The K-step loop is repeated back to back (no synchronisation in between) until the timed region lasts --min-seconds
(default 1 s): `steps` is K as asked, `repetitions` says how often it ran, `value` is over the whole region, so the
number is a sustained one and is divided by the sustained peak.

JSON keys beyond the base contract:
  value     device-resident throughput: inputs already in HBM, CUDA events on the evaluator's stream
  e2e       the same metric through the public API (Evaluator.predict_into == predict_on_batch semantics) with
            pinned HOST buffers: H2D of the planes and D2H of policy+value inside the timed region
  roofline  the dominant kernel (tower_kernel: stem + all 3x3 implicit-GEMM convolutions + head 1x1 convolutions in one
            launch; with CAZ_NO_TOWER=1 the per-layer conv kernel): algorithmic FLOPs per launch / its mean launch time,
            measured live with CUDA events around the tower stage of every timed step; frac_burst / frac_sustained
  small_batch  roofline block of the batch-16 persistent kernel: algorithmic bytes, L2->SM bytes, latency
  selfplay  BASELINE configs[2] (N=1) / configs[3] (N>1, games sharded by rank, summed over ranks)
  cpu_baseline  the oracle (torch-CPU restatement of the Keras graph) on this box's host cores, bounded sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "positions/sec (7x256, batch 4096)"
UNIT = "positions/s"


def flops_per_position(blocks=7, filters=256, stem_k=5, planes=18, pf=2, vf=4, vfc=256, labels=1968):
    """2*Cin*k^2*Cout*64 per conv (zero-padded taps counted), Dense 2*in*out."""
    macs = planes * stem_k * stem_k * filters * 64 + 2 * blocks * filters * 9 * filters * 64
    macs += filters * pf * 64 + filters * vf * 64 + pf * 64 * labels + vf * 64 * vfc + vfc
    return 2 * macs


def conv3x3_flops_per_launch(batch, filters=256):
    return 2 * 64 * batch * filters * 9 * filters


def ncu_facts():
    """Numbers that only Nsight Compute can give (DRAM bytes per launch of the dominant kernel, L2->SM bytes per pass of
    the small-batch kernel), copied by hand from the newest capture under profiles/ into profiles/roofline_traffic.json
    together with the name of that capture."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def tower_flops_per_launch(batch, blocks=7, filters=256, stem_k=5, planes=18, pf=2, vf=4):
    """Algorithmic FLOPs of ONE tower_kernel launch: stem + 2*blocks 3x3 convolutions + the heads' 1x1 convolutions."""
    macs = planes * stem_k * stem_k * filters * 64 + 2 * blocks * filters * 9 * filters * 64 + filters * (pf + vf) * 64
    return 2 * macs * batch


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def synthetic_batch(batch, seed, ev):
    """Seeded legal-looking positions, encoded by the product itself: FEN -> 68-byte records on the host, records ->
    input planes by the encoder kernel of evaluator `ev` (planes are what predict_on_batch receives)."""
    from chess_alpha_zero_b200 import encoder
    base = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(256, seed=seed)))
    reps = (batch + 255) // 256
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1, 1))[:batch])


def synthetic_batch_oracle(batch, seed):
    """The same kind of input for the CPU-only reference arm (no GPU in that process): oracle generator and encoder."""
    from oracle import encoder as oenc
    base = np.stack([oenc.canon_input_planes(f) for f in oenc.synthetic_fens(256, seed=seed)])
    reps = (batch + 255) // 256
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1, 1))[:batch])


def network(blocks):
    """Random-init weights of the named architecture (Keras default initialisers), seeded."""
    from chess_alpha_zero_b200 import keras_graph
    cfg = keras_graph.build_config(blocks=blocks)
    return cfg, keras_graph.init_weights(cfg, seed=0)


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        import torch
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        # NCCL writes its banner ("NCCL version ...") to stdout when the communicator comes up; stdout carries exactly
        # one JSON line, so the rendezvous and a first barrier run with fd 1 pointed at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
            dist.barrier()
            if backend == "nccl":
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return world, rank, local


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def cpu_baseline(cfg, weights, planes, seconds=12.0):
    """Oracle (port of the Keras graph, torch CPU fp32, all host cores) on a bounded sample of the workload."""
    import torch
    from oracle import network as onet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = planes[:256]
    onet.forward_graph(cfg, weights, sample[:32])
    best, t_end, reps = None, time.perf_counter() + seconds, 0
    while time.perf_counter() < t_end or reps < 2:
        t0 = time.perf_counter()
        onet.forward_graph(cfg, weights, sample)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
    return {"value": len(sample) / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(sample)} positions of the batch-4096 workload per pass, best of {reps} passes, "
                      f"oracle.network.forward_graph (torch CPU fp32, {cores} threads)"}


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (Keras/TF are not
    installable offline), all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import network as onet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, weights = network(args.blocks)
    # the full batch per step if the whole run then ends within ~2.5 minutes on this box, else the largest multiple of 256
    probe = synthetic_batch_oracle(256, seed=11)
    onet.forward_graph(cfg, weights, probe[:64])
    t0 = time.perf_counter()
    onet.forward_graph(cfg, weights, probe)
    rate = 256 / (time.perf_counter() - t0)
    budget = 150.0 * rate / max(1, args.steps + args.warmup)
    sample_n = int(min(args.batch, max(256, budget // 256 * 256)))
    planes = synthetic_batch_oracle(sample_n, seed=11)
    for _ in range(args.warmup):
        onet.forward_graph(cfg, weights, planes)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        onet.forward_graph(cfg, weights, planes)
    dt = time.perf_counter() - t0
    value = sample_n * args.steps / dt
    sample = (f"{sample_n} positions per step (bounded sample of the batch-{args.batch} workload), "
              f"oracle.network.forward_graph, torch CPU fp32, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.blocks}x256 policy/value net, 5x5 stem, batch {args.batch} per GPU "
                               f"(BASELINE.json configs[{1 if args.blocks == 7 else 4}])",
                   "res_blocks": args.blocks, "batch_per_gpu": args.batch, "positions_per_step": sample_n,
                   "note": "CPU arm: one process on rank 0, all host threads" +
                           ("" if sample_n == args.batch else f"; {sample_n} of the {args.batch} positions per step "
                            "(bounded sample: positions/s does not depend on the batch at this size)")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def training_leg(args, cfg, weights, ev, local, rank, world, batch=384):
    """SURVEY.md 8 f4 (not the headline metric): the `opt` worker's step -- training-mode BatchNormalization, cross-entropy x 1.25 +
    MSE, l2, Adam (worker/optimize.py:79-103) -- at batch 384 per GPU on this net, tensor-core build.  With N > 1 every rank
    trains on its own shard and the flat gradient vector is summed with torch.distributed.all_reduce (NCCL): the one collective
    of this repository.  Returns positions/s over all ranks (host-visible wall time, max over ranks)."""
    import torch
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import encoder
    planes = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(batch, seed=300 + rank)))
    rng = np.random.RandomState(rank)
    pt = np.zeros((batch, ev.n_labels), np.float32)
    for i in range(batch):
        idx = rng.permutation(ev.n_labels)[:30]
        pt[i, idx] = rng.dirichlet([0.5] * 30)
    vt = rng.uniform(-1, 1, size=batch).astype(np.float32)
    tr = caz.Trainer(cfg, weights, device=local, precision="bf16", max_batch=batch)
    losses = [tr.train_on_batch(planes, pt, vt)["loss"] for _ in range(3)]
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.train_steps):
        losses.append(tr.train_on_batch(planes, pt, vt)["loss"])
    dt = max_over_ranks(time.perf_counter() - t0, world)
    digest = float(np.abs(np.concatenate([v.reshape(-1) for v in tr.weights().values()])).sum())
    same = max_over_ranks(digest, world) == -max_over_ranks(-digest, world)
    tr.close()
    return {"metric": "training positions/s (forward + backward + Adam, batch 384 per GPU)", "value": world * batch * args.train_steps / dt,
            "unit": UNIT, "ms_per_step": 1e3 * dt / args.train_steps, "steps": args.train_steps, "n_gpus": world, "precision": "bf16 (mixed)",
            "approx_tflops": 3 * flops_per_position(args.blocks) * world * batch * args.train_steps / dt / 1e12,
            "loss_first_last": [losses[0], losses[-1]], "ranks_hold_identical_weights": bool(same),
            "collective": ("torch.distributed.all_reduce (NCCL) of the flat fp32 gradient vector, every step; the upper half of the "
                           "network's gradients is reduced under the backward pass of the lower half") if world > 1 else None}


def latency_probe(ev, planes16, calls=1000):
    """p50 of the host-visible call at batch 16 (pinned buffers), and of the device part (CUDA events)."""
    import torch
    from chess_alpha_zero_b200 import _lib
    pin_x, pin_p, pin_v = _lib.PinnedArray((16, 18, 8, 8), np.float32), _lib.PinnedArray((16, 1968), np.float32), \
        _lib.PinnedArray((16, 1), np.float32)
    pin_x.array[:] = planes16
    for _ in range(50):
        ev.predict_into(pin_x.array, pin_p.array, pin_v.array)
    host = []
    for _ in range(calls):
        t0 = time.perf_counter()
        ev.predict_into(pin_x.array, pin_p.array, pin_v.array)
        host.append(time.perf_counter() - t0)
    # the same call from 68-byte board records, the form the batched search submits (the encoder runs inside the kernel, which
    # reads the records straight from the pinned buffer)
    from chess_alpha_zero_b200 import encoder
    pin_r = _lib.PinnedArray((16, _lib.RECORD_BYTES), np.uint8)
    pin_r.array[:] = encoder.fens_to_records(encoder.synthetic_fens(16, seed=3))
    for _ in range(50):
        ev.predict_records_into(pin_r.array, pin_p.array, pin_v.array)
    host_rec = []
    for _ in range(calls):
        t0 = time.perf_counter()
        ev.predict_records_into(pin_r.array, pin_p.array, pin_v.array)
        host_rec.append(time.perf_counter() - t0)
    stream = torch.cuda.ExternalStream(ev.stream)
    d_x = torch.from_numpy(planes16).cuda()
    d_p = torch.empty((16, 1968), dtype=torch.float32, device="cuda")
    d_v = torch.empty((16,), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    dev = []
    for _ in range(min(calls, 300)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ev.eval_device(d_x.data_ptr(), 16, d_p.data_ptr(), d_v.data_ptr())
        e1.record(stream)
        e1.synchronize()
        dev.append(e0.elapsed_time(e1) * 1e-3)
    # the same launches timed by the library's own CUDA events, recorded right before and after the launch inside
    # caz_eval_device (the torch events above also contain the Python time between e0.record and the launch)
    lib_dev = []
    ev.profile(True)
    ev.profile_read(reset=True)
    for _ in range(min(calls, 300)):
        ev.eval_device(d_x.data_ptr(), 16, d_p.data_ptr(), d_v.data_ptr())
        st = ev.profile_read(reset=True)
        lib_dev.append(sum(v["ms"] for v in st.values() if v["launches"] > 0) * 1e-3)
    ev.profile(False)
    q = lambda a, f: float(np.quantile(np.asarray(a), f) * 1e6)
    return {"batch": 16, "calls": calls, "host_call_p50_us": q(host, 0.5), "host_call_p99_us": q(host, 0.99),
            "host_call_records_p50_us": q(host_rec, 0.5), "host_call_records_p99_us": q(host_rec, 0.99),
            "device_p50_us": q(lib_dev, 0.5), "device_p99_us": q(lib_dev, 0.99),
            "device_incl_python_launch_p50_us": q(dev, 0.5), "device_incl_python_launch_p99_us": q(dev, 0.99),
            "note": "host call = predict_on_batch-equivalent with pinned buffers incl. H2D+D2H+sync; host_call_records = the same from "
                    "68-byte board records (caz_eval_records, no input copy); pipe round trips excluded; "
                    "device = CUDA events recorded by the library around its launch (launch overhead included)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="positions per GPU per step")
    ap.add_argument("--blocks", type=int, default=7, help="residual blocks (19 = config 5)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16_pure", "fp16", "fp32"])
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="repeat the K-step loop back to back until the timed region is at least this long")
    ap.add_argument("--train-steps", type=int, default=10,
                    help="SURVEY 8 f4: timed steps of the training leg (forward + backward + Adam at batch 384 per GPU, gradients "
                         "all-reduced with NCCL when N > 1; 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--selfplay-seconds", type=float, default=8.0,
                    help="BASELINE configs[2] / [3]: seconds of 256-game / 800-sim self-play per GPU through this evaluator "
                         "(0 = skip; N>1: every rank plays its own games, sims/s summed over ranks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world, rank, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import torch
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    B = args.batch
    cfg, weights = network(args.blocks)
    ev = caz.Evaluator(cfg, weights, device=local, precision=args.precision, max_batch=B)
    stream = torch.cuda.ExternalStream(ev.stream)

    # inputs: NBUF distinct device batches (> L2 together with the 0.7 GB of activations a step streams)
    NBUF = 8
    host_batches = [synthetic_batch(B, seed=100 + rank * NBUF + i, ev=ev) for i in range(2)]
    d_in = [torch.from_numpy(host_batches[i % 2]).cuda() for i in range(NBUF)]
    d_pol = torch.empty((B, ev.n_labels), dtype=torch.float32, device="cuda")
    d_val = torch.empty((B,), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def step(i):
        ev.eval_device(d_in[i % NBUF].data_ptr(), B, d_pol.data_ptr(), d_val.data_ptr())

    for i in range(args.warmup):
        step(i)
    ev.sync()
    # how often the K-step loop must repeat for the timed region to last --min-seconds (the same count on every rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(3):
        step(i)
    e1.record(stream)
    e1.synchronize()
    est_step_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3 / 3, world)
    reps = max(1, int(np.ceil(args.min_seconds / max(est_step_s * args.steps, 1e-9))))
    ev.profile(True)
    ev.profile_read(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier(world)
    torch.cuda.synchronize()
    launches0 = ev.kernel_launches
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    t_wall0 = time.perf_counter()
    marks[0].record(stream)
    for r in range(reps):
        for i in range(args.steps):
            step(r * args.steps + i)
        marks[r + 1].record(stream)
    marks[-1].synchronize()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier(world)
    launches = ev.kernel_launches - launches0
    stages = ev.profile_read(reset=True)
    ev.profile(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    timed_steps = reps * args.steps
    dev_s = max_over_ranks(marks[0].elapsed_time(marks[-1]) * 1e-3, world)
    first_rep_s = max_over_ranks(marks[0].elapsed_time(marks[1]) * 1e-3, world)
    value = world * B * timed_steps / dev_s

    # ---- e2e: host buffers in, host buffers out, through the public API -----------------------------
    # Two batches in flight (Evaluator.submit / collect = caz_submit / caz_collect): every step's planes are copied
    # from pinned host memory and every step's policy + value are copied back and read on the host, inside the timed
    # region; step i+1's H2D and step i-1's D2H run under step i's forward pass.
    pins = [(_lib.PinnedArray((B, 18, 8, 8), np.float32), _lib.PinnedArray((B, ev.n_labels), np.float32),
             _lib.PinnedArray((B, 1), np.float32)) for _ in range(2)]
    for i, (px, _, _) in enumerate(pins):
        px.array[:] = host_batches[i % 2]

    def e2e_loop(n):
        tickets = [None, None]
        acc = 0.0
        for i in range(n):
            s_ = i & 1
            if tickets[s_] is not None:
                ev.collect(tickets[s_])
                acc += float(pins[s_][2].array[0, 0])      # the step's result is really read on the host
            px, pp, pv = pins[s_]
            tickets[s_] = ev.submit(px.array, pp.array, pv.array)
        for s_ in (0, 1):
            if tickets[s_] is not None:
                ev.collect(tickets[s_])
                acc += float(pins[s_][2].array[0, 0])
        return acc

    e2e_loop(4)
    barrier(world)
    t0 = time.perf_counter()
    loss_like = e2e_loop(timed_steps)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    barrier(world)
    # the plain synchronous call, for comparison (what ChessModelAPI's worker thread does per batch)
    px, pp, pv = pins[0]
    t0 = time.perf_counter()
    for _ in range(max(args.steps // 4, 5)):
        ev.predict_into(px.array, pp.array, pv.array)
    sync_s = (time.perf_counter() - t0) / max(args.steps // 4, 5)
    e2e = {"value": world * B * timed_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 18 * 64 * 4,
           "d2h_bytes_per_step": B * (ev.n_labels + 1) * 4, "ms_per_step": 1e3 * e2e_s / timed_steps,
           "steps_timed": timed_steps,
           "api": "Evaluator.submit/collect (caz_submit/caz_collect: predict_on_batch semantics, pinned host buffers, "
                  "two batches in flight)", "sync_call_positions_per_s": B / sync_s,
           "sync_call_api": "Evaluator.predict_into (caz_eval), one batch at a time", "checksum": loss_like}

    selfplay = None
    if args.selfplay_seconds > 0 and args.batch >= 2048:
        # configs[2] (one GPU) / configs[3] (N GPUs: every rank plays its own 256 games on its own evaluator, no collective):
        # the host-side batched search (include/caz_mcts.h) feeding caz_submit_records, game groups alternating
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import selfplay_bench
        barrier(world)
        mine = selfplay_bench.run(ev, games=256, sims=800, search_threads=16, seconds=args.selfplay_seconds, groups=0 if world > 1 else 2,
                                  rank=rank, world=world, seed=1 + rank,
                                  label=f"{args.blocks}x256 net ({args.precision}), random-init weights")
        if world == 1:
            selfplay = mine
        else:
            slowest = max_over_ranks(mine["seconds"], world)
            selfplay = {"workload": mine["workload"] + f", per GPU x {world} GPUs (games sharded by rank, no collective)",
                        "n_gpus": world, "seconds": slowest,
                        "sims_per_s": sum_over_ranks(mine["sims_per_s"] * mine["seconds"], world) / slowest,
                        "positions_per_s_through_evaluator":
                            sum_over_ranks(mine["positions_per_s_through_evaluator"] * mine["seconds"], world) / slowest,
                        "evaluator_occupancy_min": -max_over_ranks(-mine["evaluator_occupancy"], world),
                        "evaluator_occupancy_mean": sum_over_ranks(mine["evaluator_occupancy"], world) / world,
                        "moves_played": int(sum_over_ranks(mine["moves_played"], world)),
                        "games_per_hour_at_80_moves": sum_over_ranks(mine["games_per_hour_at_80_moves"], world),
                        "groups": mine["groups"], "host_threads_per_group": mine["host_threads_per_group"],
                        "host_cores": mine["host_cores"], "rank0": mine}
    training = None
    if args.train_steps > 0:
        training = training_leg(args, cfg, weights, ev, local, rank, world)
    if rank != 0:
        return
    peaks = measured_peaks()
    tower = stages["tower"]
    n_launch = max(tower["launches"], 1)
    fused = tower["launches"] <= timed_steps          # one tower_kernel launch per step (stem inside it)
    launch_ms = tower["ms"] / n_launch
    flops_launch = tower_flops_per_launch(B, args.blocks) if fused else conv3x3_flops_per_launch(B)
    achieved = flops_launch / (launch_ms * 1e-3) / 1e12
    burst, sustained = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    long_region = dev_s >= 1.0                        # a seconds-long region runs under the power cap: sustained peak
    peak = sustained if long_region else burst
    facts = ncu_facts()
    traffic = facts.get("tower_kernel" if fused else "conv_tc2_kernel", {})
    roofline = {"bound": "tensor",
                "kernel": "tower_kernel (stem + all 3x3x256->256 implicit-GEMM convolutions + head 1x1 convolutions, one launch)"
                          if fused else "conv_tc2_kernel (one 3x3x256->256 implicit-GEMM convolution per launch)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "frac_burst": achieved / burst, "frac_sustained": achieved / sustained,
                "peak_source": f"{peaks['source']}: {'bf16_tflops_sustained' if long_region else 'bf16_tflops (burst)'} because the "
                               f"timed region lasted {dev_s:.2f} s; burst {burst}, sustained {sustained}",
                "traffic": traffic.get("dram_bytes_per_launch") if traffic.get("batch") == B and traffic.get("blocks", 7) == args.blocks else None,
                "traffic_source": traffic.get("source"),
                "flops_per_launch": flops_launch, "mean_launch_ms": launch_ms, "launches_timed": n_launch,
                "whole_net_tflops": flops_per_position(args.blocks) * value / world / 1e12,
                "whole_net_frac_of_peak": flops_per_position(args.blocks) * value / world / 1e12 / peak,
                "whole_net_frac_burst": flops_per_position(args.blocks) * value / world / 1e12 / burst,
                "stage_ms_per_step": {k: v["ms"] / timed_steps for k, v in stages.items()}}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / timed_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "repetitions": reps, "timed_steps": timed_steps, "timed_region_s": dev_s,
        "first_repetition": {"steps": args.steps, "ms_per_step": 1e3 * first_rep_s / args.steps,
                             "value": world * B * args.steps / first_rep_s},
        "dtype": {"bf16": "bf16", "bf16_pure": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": f"{args.blocks}x256 policy/value net, 5x5 stem, batch {B} per GPU "
                               f"(BASELINE.json configs[{1 if args.blocks == 7 else 4}])",
                   "res_blocks": args.blocks, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"replicas x{world}, positions sharded, no collective",
                   "weights": "random-init (Keras default initialisers), seeded",
                   "operands": {"bf16": "weights bf16, activation operand copies fp16 (saturating), fp32 accumulation in TMEM, "
                                        "fp32 residual stream and heads", "bf16_pure": "weights and activation operands bf16",
                                "fp16": "weights and activation operands fp16", "fp32": "fp32 CUDA cores"}[args.precision],
                   "timed_region": f"{reps} back-to-back repetitions of the {args.steps}-step loop = {timed_steps} steps, "
                                   f"{dev_s:.2f} s (--min-seconds {args.min_seconds})",
                   "l2": f"inputs rotate over {NBUF} device batches; a step streams ~{1.9e-3 * B:.1f} GB of "
                         "activations through HBM (>> 126 MB L2); the 17.4 MB weight set is meant to stay L2-resident"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "flops_per_position": flops_per_position(args.blocks), "host_cores": os.cpu_count(),
    }
    if not args.no_latency and world == 1:
        lat = latency_probe(ev, host_batches[0][:16].copy())
        out["latency_b16"] = lat
        # roofline block of the batch-16 persistent kernel: what it must read per pass (the folded weights, L2-resident,
        # plus 12.5 KB per position in and out) against the time the device takes; L2->SM bytes from the newest ncu capture
        alg = ev.weight_bytes + 16 * (18 * 64 * 4 + (ev.n_labels + 1) * 4)
        sb = facts.get("forward_sb_kernel", {})
        out["small_batch"] = {"batch": 16, "kernel": "forward_sb_kernel<pair> (whole forward pass, one cooperative launch)",
                              "bound": "latency chain over L2 (15 layers x [MMA phase at the shared-memory roofline + fence/flag/poll/TMA])",
                              "algorithmic_bytes": alg, "device_p50_us": lat["device_p50_us"],
                              "achieved_gbs": alg / (lat["device_p50_us"] * 1e-6) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
                              "frac_of_hbm_peak": alg / (lat["device_p50_us"] * 1e-6) / 1e9 / peaks["hbm_gbs"],
                              "l2_to_sm_bytes_per_pass": sb.get("l2_to_sm_bytes_per_pass"),
                              "l2_to_sm_gbs": (sb["l2_to_sm_bytes_per_pass"] / (lat["device_p50_us"] * 1e-6) / 1e9)
                              if sb.get("l2_to_sm_bytes_per_pass") else None,
                              "dram_bytes_per_pass": sb.get("dram_bytes_per_pass"), "ncu_source": sb.get("source")}
    if selfplay is not None:
        out["selfplay"] = selfplay
    if training is not None:
        out["training"] = training
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(cfg, weights, host_batches[0])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
