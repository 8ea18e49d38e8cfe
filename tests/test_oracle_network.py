"""Network oracle self-consistency (PARITY UNPINNED against Keras/TF, see oracle/__init__.py):
two independent formulations must agree, and the config generator must reproduce the layer list
of the one Keras JSON the reference ships.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import encoder, network, weights

SHIPPED = "/root/reference/data/model/model_best_config.json"


def _planes(n, seed=5):
    return np.stack([encoder.canon_input_planes(f) for f in encoder.synthetic_fens(n, seed)])


@pytest.mark.skipif(not os.path.exists(SHIPPED), reason="reference checkout absent (GPU box)")
def test_config_generator_reproduces_shipped_json():
    shipped = json.load(open(SHIPPED))
    assert weights.keras_config(stem_k=3, value_filters=1) == shipped


@pytest.mark.parametrize("recipe", ["keras_default", "randomized"])
@pytest.mark.parametrize("variant", [dict(stem_k=5, value_filters=4), dict(stem_k=3, value_filters=1)])
def test_graph_interpreter_agrees_with_folded_gemm(recipe, variant):
    cfg = weights.keras_config(blocks=2, filters=64, **variant)
    w = weights.synth_weights(cfg, seed=3, recipe=recipe)
    x = _planes(6)
    p1, v1 = network.forward_graph(cfg, w, x)
    p2, v2 = network.forward_folded(cfg, w, x)
    assert p1.shape == (6, 1968) and v1.shape == (6, 1)
    assert np.abs(p1 - p2).max() <= 1e-5 and np.abs(v1 - v2).max() <= 1e-5
    assert np.allclose(p1.sum(axis=1), 1, atol=1e-5) and np.all(np.abs(v1) <= 1)


def test_full_size_net_counts_and_agreement():
    cfg = weights.keras_config()                       # 7 x 256, 5x5 stem, 4-filter value conv
    w = weights.synth_weights(cfg, seed=0, recipe="randomized")
    assert sum(v.size for v in w.values()) == 8_709_577
    x = _planes(3)
    p1, v1 = network.forward_graph(cfg, w, x)
    p2, v2 = network.forward_folded(cfg, w, x)
    assert np.abs(p1 - p2).max() <= 1e-5 and np.abs(v1 - v2).max() <= 1e-5
    pb, vb = network.forward_bf16(cfg, w, x)
    assert np.abs(pb - p1).max() <= 2e-2 and np.abs(vb - v1).max() <= 2e-2
    assert np.abs(v1).max() < 0.98                      # the randomized recipe keeps tanh unsaturated
