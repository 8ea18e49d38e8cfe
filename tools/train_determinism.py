"""Run-to-run spread of the training step's gradient vector: the same positions through caz_train_forward_backward
several times from the same weights, per tensor, both builds.  (The step has atomicAdd reductions; this shows where their
rounding noise ends up.)

    python tools/train_determinism.py [--blocks 7 --filters 256 --batch 384 --repeats 4]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=7)
    ap.add_argument("--filters", type=int, default=256)
    ap.add_argument("--batch", type=int, default=384)
    ap.add_argument("--repeats", type=int, default=4)
    args = ap.parse_args()
    import chess_alpha_zero_b200 as caz
    from chess_alpha_zero_b200 import encoder, keras_graph
    cfg = keras_graph.build_config(blocks=args.blocks, filters=args.filters)
    w = keras_graph.init_weights(cfg, seed=0)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=512)
    B = args.batch
    planes = ev.encode(encoder.fens_to_records(encoder.synthetic_fens(B, seed=50)))
    ev.close()
    rng = np.random.RandomState(0)
    pt = np.zeros((B, 1968), np.float32)
    for i in range(B):
        idx = rng.permutation(1968)[:30]
        pt[i, idx] = rng.dirichlet([0.5] * 30)
    vt = rng.uniform(-1, 1, size=B).astype(np.float32)
    for precision in ("fp32", "bf16"):
        tr = caz.Trainer(cfg, w, precision=precision, max_batch=B, distributed=False)
        runs = []
        for _ in range(args.repeats):
            tr.forward_backward(planes, pt, vt)
            runs.append(tr.gradients())
        outs = tr.outputs(B)
        tr.forward_backward(planes, pt, vt)
        outs2 = tr.outputs(B)
        tr.close()
        worst = {}
        for name in runs[0]:
            ref = runs[0][name]
            d = max(float(np.abs(r[name] - ref).max()) for r in runs[1:])
            worst[name] = d / max(float(np.abs(ref).max()), 1e-30)
        top = sorted(worst.items(), key=lambda kv: -kv[1])[:8]
        print(json.dumps({"precision": precision, "net": f"{args.blocks}x{args.filters}", "batch": B,
                          "forward_outputs_identical": bool(np.array_equal(outs[0], outs2[0]) and np.array_equal(outs[1], outs2[1])),
                          "tensors_bit_identical": sum(v == 0 for v in worst.values()), "tensors": len(worst),
                          "largest_relative_spread": {k: float(f"{v:.3g}") for k, v in top}}))


if __name__ == "__main__":
    main()
