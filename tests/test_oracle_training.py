"""The training oracle (oracle/training.py) against itself: autograd gradients vs central finite differences of its own
loss in float64, BatchNormalization moving averages, Adam's first step.  CPU only."""
import numpy as np
import torch

from oracle import encoder as oenc
from oracle import training as otrain
from oracle import weights as oweights


def _tiny():
    cfg = oweights.keras_config(filters=8, blocks=1, value_fc=16)
    w = oweights.synth_weights(cfg, seed=3, recipe="randomized")
    fens = oenc.synthetic_fens(6, seed=4)
    x = np.stack([oenc.canon_input_planes(f) for f in fens])
    x[:, 16] /= 50.0                                              # keep the clock plane from saturating the tiny net
    rng = np.random.RandomState(0)
    pt = rng.dirichlet([0.05] * 1968, size=6).astype(np.float32)
    vt = rng.uniform(-1, 1, size=6).astype(np.float32)
    return cfg, w, x, pt, vt


def test_gradients_match_finite_differences():
    cfg, w, x, pt, vt = _tiny()
    tr = otrain.OracleTrainer(cfg, w)
    out, g, data, _, _ = tr.gradients(x, pt, vt)
    assert out["total"] > 0 and abs(out["total"] - (1.25 * out["ce"] + out["mse"] + out["l2"])) < 1e-12
    rng = np.random.RandomState(1)

    def loss_at(name, idx, delta):
        ww = {k: v.clone() for k, v in tr.w.items()}
        ww[name].view(-1)[idx] += delta
        p, v, _ = otrain.forward_train(cfg, ww, torch.from_numpy(x).double())
        return float(otrain.losses(cfg, ww, p, v, torch.from_numpy(pt).double(), torch.from_numpy(vt).double())[0])

    for name in ("input_conv-5-8/kernel:0", "res1_conv2-3-8/kernel:0", "res1_batchnorm1/gamma:0", "policy_out/kernel:0",
                 "value_dense/bias:0", "value_out/kernel:0", "input_batchnorm/beta:0"):
        flat = g[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=3, replace=False):
            h = 1e-5
            fd = (loss_at(name, int(idx), h) - loss_at(name, int(idx), -h)) / (2 * h)
            assert abs(fd - float(flat[idx])) <= 1e-6 + 1e-5 * abs(fd), (name, idx, fd, float(flat[idx]))
    assert all(np.allclose(data[k].numpy(), (g[k] - 2e-4 * tr.w[k]).numpy()) for k in tr.kernels)


def test_first_adam_step_and_moving_averages():
    cfg, w, x, pt, vt = _tiny()
    tr = otrain.OracleTrainer(cfg, w)
    _, g, _, _, stats = tr.gradients(x, pt, vt)
    before = {k: v.clone() for k, v in tr.w.items()}
    tr.train_on_batch(x, pt, vt)
    # t = 1: m = 0.1 g, v = 0.001 g^2, lr_t = lr * sqrt(0.001) / 0.1  =>  step = lr * g / (|g| + eps / sqrt(0.001))
    k = "res1_conv1-3-8/kernel:0"
    step = (before[k] - tr.w[k]).numpy()
    gk = g[k].numpy()
    assert np.allclose(step, 1e-3 * gk / (np.abs(gk) + 1e-8 / np.sqrt(1e-3)), rtol=1e-9, atol=1e-15)
    mm = "input_batchnorm/moving_mean:0"
    assert np.allclose(tr.w[mm].numpy(), 0.99 * before[mm].numpy() + 0.01 * stats[mm].numpy())
    mv = "res1_batchnorm2/moving_variance:0"
    assert np.allclose(tr.w[mv].numpy(), 0.99 * before[mv].numpy() + 0.01 * stats[mv].numpy())
    assert tr.step == 1
