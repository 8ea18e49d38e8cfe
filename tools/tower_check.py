"""First-contact check of the one-launch tower (conv_tower.cuh) against the per-layer chain: bit-identity and time per batch.
    python tools/tower_check.py [batches...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_alpha_zero_b200 as caz  # noqa: E402
from chess_alpha_zero_b200 import encoder, keras_graph  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [5, 100, 300, 512]
    blocks = int(os.environ.get("BLOCKS", "7"))
    cfg = keras_graph.build_config(blocks=blocks)
    w = keras_graph.init_weights(cfg, seed=0)
    cap = max(batches)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=cap)
    ev.set_small_batch_limit(0)
    base = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(min(cap, 512), seed=3)))
    x = np.ascontiguousarray(np.tile(base, ((cap + len(base) - 1) // len(base), 1, 1, 1))[:cap])
    ok = True
    for b in batches:
        ev.set_fused_tower(0)
        want = ev.predict_on_batch(x[:b])
        row = {"batch": b}
        for mode in (1, 2, -1):
            ev.set_fused_tower(mode)
            got = ev.predict_on_batch(x[:b])
            same = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
            row[f"mode{mode}_identical"] = same
            if not same:
                row[f"mode{mode}_max_dp"] = float(np.abs(got[0] - want[0]).max())
                row[f"mode{mode}_max_dv"] = float(np.abs(got[1] - want[1]).max())
            ok &= same
        for mode in (0, 1, 2):
            ev.set_fused_tower(mode)
            for _ in range(5):
                ev.predict_on_batch(x[:b])
            ev.profile(True)
            ev.profile_read(reset=True)
            n = 30
            for _ in range(n):
                ev.predict_on_batch(x[:b])
            st = ev.profile_read(reset=True)
            ev.profile(False)
            row[f"mode{mode}_device_us"] = round(sum(v["ms"] for v in st.values()) / n * 1e3, 1)
            row[f"mode{mode}_tower_us"] = round((st["stem"]["ms"] + st["tower"]["ms"]) / n * 1e3, 1)
        print(json.dumps(row), flush=True)
    ev.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
