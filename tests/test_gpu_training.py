"""The GPU training step (include/caz_b200.h: caz_train_*) against the CPU oracle (oracle/training.py:
torch autograd in float64, Keras semantics): losses, every gradient tensor, Adam updates and BatchNormalization moving
averages over several steps.  Needs a B200 (-m gpu)."""
import numpy as np
import pytest

import chess_alpha_zero_b200 as caz
from oracle import encoder as oenc
from oracle import training as otrain
from oracle import weights as oweights

pytestmark = pytest.mark.gpu


def _data(n, seed):
    fens = oenc.synthetic_fens(n, seed=seed)
    x = np.stack([oenc.canon_input_planes(f) for f in fens])
    rng = np.random.RandomState(seed)
    pt = np.zeros((n, 1968), np.float32)
    for i in range(n):                                               # visit-count-like targets on 5..40 moves
        k = rng.randint(5, 41)
        idx = rng.permutation(1968)[:k]
        pt[i, idx] = rng.dirichlet([0.5] * k)
    vt = rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32) * rng.uniform(0.2, 1.0, size=n).astype(np.float32)
    return x, pt, vt


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("variant", [dict(filters=64, blocks=2), dict(filters=32, blocks=1, stem_k=3, value_filters=1)])
def test_gradients_and_losses_match_the_oracle(variant):
    cfg = oweights.keras_config(**variant)
    w = oweights.synth_weights(cfg, seed=2, recipe="randomized")
    x, pt, vt = _data(24, seed=5)
    ora = otrain.OracleTrainer(cfg, w)
    want, _, want_g, (wp, wv), _ = ora.gradients(x, pt, vt)
    tr = caz.Trainer(cfg, w, precision="fp32", max_batch=32)
    got = tr.forward_backward(x, pt, vt)
    assert abs(got["loss"] - want["total"]) <= 1e-4 * want["total"]
    assert abs(got["policy_out_loss"] - want["ce"]) <= 1e-4 * want["ce"] and abs(got["value_out_loss"] - want["mse"]) <= 1e-4
    assert abs(got["l2"] - want["l2"]) <= 1e-5 * want["l2"]
    p, v = tr.outputs(24)
    assert np.abs(p - wp).max() <= 1e-5 and np.abs(v - wv).max() <= 1e-5
    g = tr.gradients()
    worst = {}
    for name, ref in want_g.items():
        worst[name] = _rel(g[name], ref.numpy())
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    print("   worst relative gradient error:", max(worst.items(), key=lambda kv: kv[1]))
    assert not bad, bad
    tr.close()


def test_adam_steps_and_moving_statistics_follow_the_oracle():
    cfg = oweights.keras_config(filters=64, blocks=2)
    w = oweights.synth_weights(cfg, seed=4, recipe="randomized")
    ora = otrain.OracleTrainer(cfg, w)
    tr = caz.Trainer(cfg, w, precision="fp32", max_batch=32)
    first = None
    for step in range(4):
        x, pt, vt = _data(16 + 4 * step, seed=20 + step)            # ragged batch sizes
        want = ora.train_on_batch(x, pt, vt)
        got = tr.train_on_batch(x, pt, vt)
        first = first or got["loss"]
        assert abs(got["loss"] - want["total"]) <= 2e-3 * want["total"], (step, got, want)
    assert tr.steps == 4
    ww, gw = ora.weights(), tr.weights()
    # Adam's first steps are ~lr * sign(g): after 4 steps weights moved by up to 4e-3; agreement to a small fraction of that
    for name in gw:
        assert np.abs(gw[name] - ww[name]).max() <= 2e-4 + 1e-3 * np.abs(ww[name] - w[name]).max(), name
    mv = "res1_batchnorm1/moving_variance:0"
    assert np.abs(gw[mv] - w[mv]).max() > 1e-4                        # the moving statistics did move
    # the trained weights load into an evaluator: Trainer.weights() -> Evaluator (ChessModel.save / caz_reload path)
    ev = caz.Evaluator(cfg, gw, precision="fp32", max_batch=16)
    p, v = ev.predict_on_batch(_data(8, seed=99)[0])
    assert p.shape == (8, 1968) and np.allclose(p.sum(1), 1, atol=1e-5)
    ev.close()
    tr.close()


def test_fit_signature_and_loss_goes_down():
    """model.fit(state_ary, [policy_ary, value_ary], batch_size, epochs, shuffle, validation_split) (optimize.py:88-93)."""
    cfg = oweights.keras_config(filters=32, blocks=1)
    w = oweights.synth_weights(cfg, seed=1, recipe="keras_default")
    x, pt, vt = _data(96, seed=7)
    x[:, 16] = 0                                                      # (the raw clock plane makes tiny nets unstable)
    tr = caz.Trainer(cfg, w, precision="fp32", max_batch=32)
    hist = tr.fit(x, [pt, vt], batch_size=32, epochs=6, shuffle=True, validation_split=0.02, seed=0)
    assert len(hist["loss"]) == 6 and hist["loss"][-1] < hist["loss"][0]
    assert tr.steps == 6 * 3                                          # 94 training samples -> 3 batches per epoch
    tr.close()


def _cos(a, b):
    a, b = a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("variant,batch", [(dict(filters=64, blocks=2), 24), (dict(filters=128, blocks=1, stem_k=3), 70)])
def test_tensor_core_build_gradients(variant, batch):
    """precision="bf16": the stem's and the tower's convolutions on tcgen05 (forward fp16 x fp16, input and weight
    gradients bf16 x bf16, fp32 accumulation), everything else fp32.  Against the float64 oracle: losses within 2e-3, every
    gradient tensor with cosine similarity >= 0.998 and within 20 % of its largest entry -- the value head's gradient is
    proportional to (v - t), so a forward error of 1e-2 on v (16-bit operands under training-mode BatchNormalization) is a
    several-percent error on everything behind it; the fp32 build is the parity build (5e-6)."""
    cfg = oweights.keras_config(**variant)
    w = oweights.synth_weights(cfg, seed=2, recipe="randomized")
    x, pt, vt = _data(batch, seed=5)
    ora = otrain.OracleTrainer(cfg, w)
    want, _, want_g, (wp, wv), _ = ora.gradients(x, pt, vt)
    tr = caz.Trainer(cfg, w, precision="bf16", max_batch=batch + 3)
    got = tr.forward_backward(x, pt, vt)
    assert abs(got["loss"] - want["total"]) <= 2e-3 * want["total"], (got, want)
    p, v = tr.outputs(batch)
    print(f"   forward: max|dp| {np.abs(p - wp).max():.2e} max|dv| {np.abs(v - wv).max():.2e}")
    assert np.abs(p - wp).max() <= 5e-3 and np.abs(v - wv).max() <= 2e-2
    g = tr.gradients()
    report = {name: (_rel(g[name], ref.numpy()), _cos(g[name], ref.numpy())) for name, ref in want_g.items()}
    worst_rel = max(report.items(), key=lambda kv: kv[1][0])
    worst_cos = min(report.items(), key=lambda kv: kv[1][1])
    print(f"   worst relative error {worst_rel[0]} {worst_rel[1][0]:.2e}; worst cosine {worst_cos[0]} {worst_cos[1][1]:.6f}")
    for name in ("input_conv", "res1_conv1", "res1_conv2", "policy_out/kernel", "value_dense/kernel"):
        k = [n for n in report if n.startswith(name) and "kernel" in n][0]
        print(f"      {k}: rel {report[k][0]:.2e} cos {report[k][1]:.6f}")
    bad = {k: v for k, v in report.items() if v[0] > 0.2 or v[1] < 0.998}
    assert not bad, bad
    tr.close()


def test_tensor_core_build_trains():
    cfg = oweights.keras_config(filters=64, blocks=2)
    w = oweights.synth_weights(cfg, seed=1, recipe="keras_default")
    x, pt, vt = _data(128, seed=7)
    x[:, 16] = 0
    tr = caz.Trainer(cfg, w, precision="bf16", max_batch=64)
    hist = tr.fit(x, [pt, vt], batch_size=64, epochs=8, shuffle=True, seed=0)
    print("   loss per epoch:", [round(v, 4) for v in hist["loss"]])
    assert hist["loss"][-1] < 0.9 * hist["loss"][0]
    ev = caz.Evaluator(cfg, tr.weights(), precision="bf16", max_batch=16)
    p, v = ev.predict_on_batch(x[:8])
    assert np.allclose(p.sum(1), 1, atol=1e-4) and np.isfinite(v).all()
    ev.close()
    tr.close()


def test_optimize_worker_calls_on_chessmodel(tmp_path):
    """The three things OptimizeWorker does with the model (worker/optimize.py:88-115): model.model.compile(...),
    model.model.fit(state_ary, [policy_ary, value_ary], batch_size, epochs, shuffle, validation_split, callbacks), model.save(...)
    -- on caz.ChessModel, whose .model is the CUDA evaluator; the saved weights load into a fresh model that predicts the same."""
    import json
    from chess_alpha_zero_b200 import keras_graph
    cfg = oweights.keras_config(filters=64, blocks=2)
    w = oweights.synth_weights(cfg, seed=1, recipe="keras_default")
    x, pt, vt = _data(96, seed=7)
    x[:, 16] = 0
    model = caz.ChessModel(caz.Config("mini"), precision="bf16", max_batch=64)
    model._install(cfg, w)
    before = model.model.predict_on_batch(x[:8])
    model.model.compile(optimizer=None, loss=["categorical_crossentropy", "mean_squared_error"], loss_weights=[1.25, 1.0])
    hist = model.model.fit(x, [pt, vt], batch_size=32, epochs=3, shuffle=True, validation_split=0.02, callbacks=[object()])
    assert len(hist.history["loss"]) == 3 and hist.history["loss"][-1] < hist.history["loss"][0]
    after = model.model.predict_on_batch(x[:8])
    assert np.abs(after[0] - before[0]).max() > 1e-6                  # the evaluator now runs the trained weights
    cpath, wpath = str(tmp_path / "model_config.json"), str(tmp_path / "model_weight.h5")
    model.save(cpath, wpath)
    fresh = caz.ChessModel(caz.Config("mini"), precision="bf16", max_batch=64)
    assert fresh.load(cpath, wpath)
    again = fresh.model.predict_on_batch(x[:8])
    assert np.array_equal(again[0], after[0]) and np.array_equal(again[1], after[1])
    model.model.close()
    fresh.model.close()


def test_gradient_halves_are_complete_when_their_events_fire():
    """caz_train_forward_backward_begin / caz_train_wait_gradients / _end, the hooks the data-parallel all-reduce hangs on:
    a side stream that waits for part 0 must see the final gradients of the upper half of the network (elements
    >= caz_train_gradient_split) although the backward pass of the lower half is still running, and after part 1 all of
    them.  Every step gets different positions, so values left by the step before cannot pass; steps 2.. replay the CUDA
    graph, where the mid-backward event is an event-record node."""
    import ctypes
    import torch
    cfg = oweights.keras_config(filters=64, blocks=3)
    w = oweights.synth_weights(cfg, seed=4, recipe="keras_default")
    B = 48
    x, pt, vt = _data(4 * B, seed=9)
    for precision in ("fp32", "bf16"):
        tr = caz.Trainer(cfg, w, precision=precision, max_batch=B, distributed=True)   # gradient vector in a torch tensor
        lib, h = tr._lib, tr._h
        split = int(lib.caz_train_gradient_split(h))
        assert 0 < split < tr.n_params
        side = torch.cuda.Stream()
        sp = ctypes.c_void_p(side.cuda_stream)
        for step in range(4):
            sl = slice(step * B, (step + 1) * B)
            xs, ps, vs = np.ascontiguousarray(x[sl]), np.ascontiguousarray(pt[sl]), np.ascontiguousarray(vt[sl])
            tr._check(lib.caz_train_forward_backward_begin(h, xs.ctypes.data, ps.ctypes.data, vs.ctypes.data, B))
            with torch.cuda.stream(side):
                tr._check(lib.caz_train_wait_gradients(h, 0, sp))
                upper = tr._grad_tensor[split:].clone()
                tr._check(lib.caz_train_wait_gradients(h, 1, sp))
                lower = tr._grad_tensor[:split].clone()
            losses = np.zeros(4, np.float32)
            tr._check(lib.caz_train_forward_backward_end(h, losses.ctypes.data))
            side.synchronize()
            final = tr._grad_tensor.clone()
            assert float(final.abs().max()) > 0 and np.isfinite(losses).all()
            assert torch.equal(upper, final[split:]), (precision, step)
            assert torch.equal(lower, final[:split]), (precision, step)
        tr.close()


def test_overlapped_all_reduce_step_runs_like_the_plain_step():
    """Trainer.train_on_batch under torch.distributed: begin, the all-reduce of the upper half of the gradient vector issued
    behind caz_train_wait_gradients(part 0), then the lower half, then Adam.  With a one-rank NCCL group the all-reduce is
    the identity, so this is the plumbing only (tools/train_bench.py --verify compares the two orders on 2..8 GPUs).  The
    tensor-core build has no atomicAdd anywhere in the step and is reproducible bit for bit: four overlapped steps must
    leave exactly the weights and losses of four plain steps.  The fp32 build's weight-gradient kernel adds its board
    slices atomically: there the agreement is to float rounding."""
    import socket
    import torch
    import torch.distributed as dist
    cfg = oweights.keras_config(filters=64, blocks=3)
    w = oweights.synth_weights(cfg, seed=4, recipe="keras_default")
    x, pt, vt = _data(32, seed=9)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        for precision in ("fp32", "bf16"):
            ta = caz.Trainer(cfg, w, precision=precision, max_batch=32, distributed=True)
            tb = caz.Trainer(cfg, w, precision=precision, max_batch=32, distributed=False)
            tc = caz.Trainer(cfg, w, precision=precision, max_batch=32, distributed=False)   # control: plain against plain
            for _ in range(4):
                la = ta.train_on_batch(x, pt, vt)
                lb = tb.train_on_batch(x, pt, vt)
                tc.train_on_batch(x, pt, vt)
                if precision == "bf16":
                    assert la == lb, (la, lb)
                assert all(abs(la[k] - lb[k]) <= 1e-5 * abs(lb[k]) for k in la), (la, lb)
            wa, wb = ta.weights(), tb.weights()
            fa = np.concatenate([v.reshape(-1) for v in wa.values()])
            fb = np.concatenate([v.reshape(-1) for v in wb.values()])
            d = np.abs(fa - fb)
            dc = np.abs(np.concatenate([v.reshape(-1) for v in tc.weights().values()]) - fb)
            tc.close()
            print(f"   {precision}: weights after 4 steps differ by max {d.max():.2e}, mean {d.mean():.2e}, {np.mean(d > 1e-5):.2e} of them > 1e-5"
                  f" (two plain runs: max {dc.max():.2e}, mean {dc.mean():.2e})")
            assert d.max() <= (0.0 if precision == "bf16" else 2e-5) and d.mean() <= 1e-8
            ta.close()
            tb.close()
    finally:
        dist.destroy_process_group()


def test_tile_kernels_and_halo_free_weight_gradient_change_nothing(monkeypatch):
    """The tensor-core build's BatchNormalization tile kernels (which also write the K-major operands of the weight-gradient
    GEMMs) against the separate elementwise + transpose kernels they replaced (CAZ_TRAIN_NO_TILE=1): with the ReLU mask read
    from the fp32 activations (CAZ_TRAIN_MASK_FP32=1) every gradient tensor and the losses are the same bit for bit.  By
    default the mask comes from the fp16 activations (half the bytes): it differs where 0 < a < 2^-25 -- no or one element of
    one layer in cases like this one -- which moves that channel's sums: bounded here at 5e-3 of each tensor's largest entry (the forward
    pass's 16-bit operands flip far more masks than that).  And the weight-gradient GEMM over the real cells only against
    the one over the whole zero-padded index (CAZ_TRAIN_WG_PADDED=1): same products, other split boundaries -> equal to
    fp32 summation rounding."""
    cfg = oweights.keras_config(filters=64, blocks=2)
    w = oweights.synth_weights(cfg, seed=2, recipe="randomized")
    x, pt, vt = _data(70, seed=5)

    def run():
        tr = caz.Trainer(cfg, w, precision="bf16", max_batch=73)
        losses = tr.forward_backward(x, pt, vt)
        g = tr.gradients()
        tr.close()
        return losses, g

    l_new, g_new = run()
    monkeypatch.setenv("CAZ_TRAIN_MASK_FP32", "1")
    l_m32, g_m32 = run()
    monkeypatch.setenv("CAZ_TRAIN_NO_TILE", "1")
    l_old, g_old = run()
    monkeypatch.delenv("CAZ_TRAIN_NO_TILE")
    monkeypatch.delenv("CAZ_TRAIN_MASK_FP32")
    assert l_new == l_old == l_m32, (l_new, l_old, l_m32)
    for k in g_old:
        assert np.array_equal(g_m32[k], g_old[k]), k
    differing = {k: _rel(g_new[k], g_old[k]) for k in g_new if not np.array_equal(g_new[k], g_old[k])}
    print(f"   fp16 mask vs fp32 mask: {len(g_new) - len(differing)} of {len(g_new)} tensors bit-identical; worst of the others "
          f"{max(differing.values(), default=0.0):.2e}")
    assert all(v <= 5e-3 for v in differing.values()), differing
    # BatchNormalization batch statistics from the convolution's epilogue (fp32 partial sums per warp, ConvArgs::stats) against
    # the separate pass over z (CAZ_TRAIN_NO_CONV_STATS=1).  The statistics themselves are visible in the gradient slots of the
    # moving averages: they agree to 2e-5.  The other tensors see the ~1e-7 shift through a few flipped 16-bit roundings -- and
    # at batch 70 ONE ReLU of the value head's hidden layer that flips for one position moves every tensor by ~1/70: cosine.
    monkeypatch.setenv("CAZ_TRAIN_NO_CONV_STATS", "1")
    l_sep, g_sep = run()
    monkeypatch.delenv("CAZ_TRAIN_NO_CONV_STATS")
    stat_names = [k for k in g_new if "moving_" in k and "policy" not in k and "value" not in k]
    assert len(stat_names) == 2 * 5
    worst_stat = max(_rel(g_new[k], g_sep[k]) for k in stat_names)
    worst_cos = min(_cos(g_new[k], g_sep[k]) for k in g_new)
    print(f"   statistics from the convolution epilogue vs a separate pass: batch mean / variance differ by {worst_stat:.2e}, "
          f"losses {l_new['loss']:.6f} / {l_sep['loss']:.6f}, worst gradient cosine {worst_cos:.6f}")
    assert worst_stat <= 2e-5 and abs(l_new["loss"] - l_sep["loss"]) <= 1e-5 * l_sep["loss"] and worst_cos >= 0.995
    monkeypatch.setenv("CAZ_TRAIN_WG_PADDED", "1")
    l_pad, g_pad = run()
    assert l_pad == l_new
    worst = max(_rel(g_new[k], g_pad[k]) for k in g_new)
    print(f"   real-cell K vs padded K: worst relative difference {worst:.2e}")
    assert worst <= 2e-5


def test_trainer_rejects_bad_calls():
    """Error behaviour of the training ABI: loud, and the trainer stays usable afterwards."""
    cfg = oweights.keras_config(filters=64, blocks=1)
    w = oweights.synth_weights(cfg, seed=1, recipe="keras_default")
    x, pt, vt = _data(8, seed=3)
    tr = caz.Trainer(cfg, w, precision="bf16", max_batch=6)
    with pytest.raises(caz.CazError, match="batch 8, max_batch 6"):
        tr.forward_backward(x, pt, vt)
    with pytest.raises(ValueError):
        tr.forward_backward(x[:4], pt[:4, :100], vt[:4])               # policy targets of the wrong width
    with pytest.raises(ValueError):
        tr.forward_backward(x[:4], pt[:4], vt[:3])                     # one value target missing
    losses = np.zeros(4, np.float32)
    rc = tr._lib.caz_train_forward_backward_end(tr._h, losses.ctypes.data)
    assert rc != 0 and b"without a begin" in tr._lib.caz_train_last_error(tr._h)
    assert tr._lib.caz_train_wait_gradients(tr._h, 2, None) != 0       # there are two parts, 0 and 1
    out = tr.train_on_batch(x[:6], pt[:6], vt[:6])                     # still works
    assert np.isfinite(out["loss"]) and tr.steps == 1
    with pytest.raises(ValueError, match="precision"):
        caz.Trainer(cfg, w, precision="fp8")
    tr.close()
    tr.close()                                                         # idempotent


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_a_short_odd_batch_after_a_long_one(precision):
    """Model.fit's last batch of an epoch is short: a step at batch 5 on a trainer that has just run batch 12 must equal the
    same step on a fresh trainer bit for bit (tensor-core build; the fp32 build to float rounding of its atomic weight-gradient
    sums) -- nothing may be read from what the longer batch left in the buffers (operand rows of boards past the batch,
    K-major operands, statistics slots)."""
    cfg = oweights.keras_config(filters=64, blocks=2)
    w = oweights.synth_weights(cfg, seed=3, recipe="keras_default")
    x, pt, vt = _data(12, seed=11)
    used = caz.Trainer(cfg, w, precision=precision, max_batch=12)
    used.forward_backward(x, pt, vt)                                   # fills every buffer with 12 boards, no update
    got_l = used.forward_backward(x[:5], pt[:5], vt[:5])
    got = used.gradients()
    used.close()
    fresh = caz.Trainer(cfg, w, precision=precision, max_batch=12)
    want_l = fresh.forward_backward(x[:5], pt[:5], vt[:5])
    want = fresh.gradients()
    fresh.close()
    if precision == "bf16":
        assert got_l == want_l
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    else:
        assert all(abs(got_l[k] - want_l[k]) <= 1e-6 * abs(want_l[k]) + 1e-9 for k in want_l)
        assert max(_rel(got[k], want[k]) for k in want) <= 1e-5
