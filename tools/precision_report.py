"""max|dp|, max|dv| and legal-masked argmax agreement of every numeric build against the fp32 oracle, over several
position sets and both summation orders (persistent small-batch kernel vs per-layer kernels).  GPU box:
    python tools/precision_report.py > gpurun_out/precision.jsonl
The CPU-side ablation that motivated the bf16 build's operand formats is profiles/r2_precision_ablation.md."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_alpha_zero_b200 as caz  # noqa: E402
from oracle import encoder as oenc  # noqa: E402
from oracle import network as onet  # noqa: E402
from oracle import weights as oweights  # noqa: E402


def main():
    nets = [("7x256 keras_default", oweights.keras_config(), "keras_default", 0),
            ("7x256 randomized", oweights.keras_config(), "randomized", 0),
            ("19x256 keras_default", oweights.keras_config(blocks=19), "keras_default", 6)]
    for tag, cfg, recipe, wseed in nets:
        w = oweights.synth_weights(cfg, seed=wseed, recipe=recipe)
        for pseed in (5, 6, 7, 8):
            fens = oenc.synthetic_fens(256, pseed)
            x = np.stack([oenc.canon_input_planes(f) for f in fens])
            wp, wv = onet.forward_graph(cfg, w, x)
            masks = onet.synthetic_legal_masks(256, seed=1)
            for prec in ("bf16", "bf16_pure", "fp16"):
                ev = caz.Evaluator(cfg, w, precision=prec, max_batch=256)
                for path in ("per-layer", "persistent"):
                    if path == "per-layer":
                        ev.set_small_batch_limit(0)
                        p, v = ev.predict_on_batch(x)
                    else:
                        ev.set_small_batch_limit(-1)
                        outs = [ev.predict_on_batch(x[i:i + 64]) for i in range(0, 256, 64)]
                        p, v = np.concatenate([o[0] for o in outs]), np.concatenate([o[1] for o in outs])
                    dv = np.abs(v - wv).reshape(-1)
                    agree = float((onet.legal_masked_argmax(p, masks) == onet.legal_masked_argmax(wp, masks)).mean())
                    print(json.dumps({"net": tag, "positions_seed": pseed, "build": prec, "path": path,
                                      "max_dp": float(np.abs(p - wp).max()), "max_dv": float(dv.max()),
                                      "rms_dv": float(np.sqrt((dv ** 2).mean())), "masked_argmax_agree": agree}), flush=True)
                ev.close()


if __name__ == "__main__":
    main()
