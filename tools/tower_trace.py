"""Per-layer timeline of the one-launch tower kernel (caz_debug_tower_trace): where a layer's time goes for pair 0.
    python tools/tower_trace.py [batch] [mode]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_alpha_zero_b200 as caz  # noqa: E402
from chess_alpha_zero_b200 import _lib, encoder, keras_graph  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cfg = keras_graph.build_config(blocks=7)
    ev = caz.Evaluator(cfg, keras_graph.init_weights(cfg, seed=0), precision="bf16", max_batch=max(batch, 16))
    ev.set_small_batch_limit(0)
    ev.set_fused_tower(mode)
    x = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(min(batch, 512), seed=3)))
    x = np.ascontiguousarray(np.tile(x, ((batch + len(x) - 1) // len(x), 1, 1, 1))[:batch])
    lib = _lib.load()
    for _ in range(5):
        ev.predict_on_batch(x)
    lib.caz_debug_tower_trace(ev._handle, 1, None, 0)
    ev.predict_on_batch(x)
    st = np.zeros(512, np.int64)
    lib.caz_debug_tower_trace(ev._handle, 0, st.ctypes.data_as(ctypes.c_void_p), 512)
    st = st.reshape(64, 8)[:15]
    t0 = st[0, 1]
    names = ["chunk0_ready", "tile_req", "chunk0_in", "mma_issued", "acc_done", "epi_done"]
    print(f"batch {batch} mode {mode}: cycles relative to the stem's first tile request (clock64 of SM 0)")
    print("layer " + " ".join(f"{n:>12s}" for n in names) + "   | ready->req  req->in  in->accdone  acc->epi_done  acc_done->next chunk0_ready  layer period")
    for l in range(15):
        r = st[l] - t0
        nxt = (st[l + 1, 0] - st[l, 4]) if l + 1 < 15 else 0
        per = (st[l + 1, 4] - st[l, 4]) if l + 1 < 15 else 0
        print(f"{l:5d} " + " ".join(f"{int(v):12d}" for v in r[:6]) +
              f"   | {int(st[l,1]-st[l,0]):9d} {int(st[l,2]-st[l,1]):8d} {int(st[l,4]-st[l,2]):12d} {int(st[l,5]-st[l,4]):14d} {int(nxt):28d} {int(per):13d}")
    ev.close()


if __name__ == "__main__":
    main()
