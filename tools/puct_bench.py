#!/usr/bin/env python
"""PUCT select (select_action_q_and_u, player_chess.py:251-299) for all concurrent games of a device: the warp-per-node
kernel on device-resident node arrays (caz_puct_select_device, CUDA events) against the host-pointer entry point
(caz_puct_select: copies + sync) and against the host-side search's own selection loop (libcaz_mcts, ns per descent level
from tools/ub/mcts_scaling).  One tree level of 256 games x 16 descents = 4 096 nodes, ~35 legal moves each."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_alpha_zero_b200 as caz  # noqa: E402
from chess_alpha_zero_b200 import _lib, keras_graph  # noqa: E402


def main():
    import torch
    cfg = keras_graph.build_config(blocks=1, filters=64)
    ev = caz.Evaluator(cfg, keras_graph.init_weights(cfg, 0), precision="bf16", max_batch=16)
    lib = _lib.load()
    rng = np.random.RandomState(0)
    for n_nodes in (256, 4096, 65536):
        sizes = rng.randint(5, 61, size=n_nodes)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        ne = int(off[-1])
        n = rng.randint(0, 50, size=ne).astype(np.int32)
        q = rng.uniform(-1, 1, size=ne)
        p = rng.dirichlet([0.3] * 8, size=ne)[:, 0]
        sum_n = np.asarray([n[off[i]:off[i + 1]].sum() for i in range(n_nodes)], np.int32)
        want = ev.puct_select(off, n, q, p, sum_n, 1.5)
        d = {k: torch.from_numpy(v).cuda() for k, v in dict(off=off, n=n, q=q, p=p, sum_n=sum_n).items()}
        best = torch.empty(n_nodes, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(ev.stream)

        def launch():
            rc = lib.caz_puct_select_device(ev._handle, d["off"].data_ptr(), n_nodes, d["n"].data_ptr(), d["q"].data_ptr(), d["p"].data_ptr(),
                                            d["sum_n"].data_ptr(), None, None, 1.5, 0.0, best.data_ptr())
            assert rc == 0
        for _ in range(10):
            launch()
        ev.sync()
        assert np.array_equal(best.cpu().numpy(), want)
        reps = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            launch()
        e1.record(stream)
        e1.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / reps
        t0 = time.perf_counter()
        for _ in range(20):
            ev.puct_select(off, n, q, p, sum_n, 1.5)
        host_api_us = (time.perf_counter() - t0) / 20 * 1e6
        t0 = time.perf_counter()
        for _ in range(20):
            launch()
            ev.sync()
            best.cpu()
        dev_roundtrip_us = (time.perf_counter() - t0) / 20 * 1e6
        bytes_in = ne * 20 + n_nodes * 8
        print(json.dumps({"nodes": n_nodes, "edges": ne, "device_kernel_us": round(dev_us, 2), "nodes_per_s_device": round(n_nodes / dev_us * 1e6),
                          "achieved_gbs": round(bytes_in / dev_us / 1e3, 1), "launch_sync_readback_us": round(dev_roundtrip_us, 1),
                          "host_pointer_entry_us": round(host_api_us, 1)}), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
