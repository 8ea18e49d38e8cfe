"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.
Needs a B200: run with ``-m gpu``.  Tolerances are north_star's: fp32 build <= 1e-4 max-abs on policy and
value; bf16 build <= 2e-2 on both with the legal-masked argmax identical on >= 99 % of positions; input
planes, prior push and PUCT select bit-exact."""
import json
import os
import threading

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import keras_graph
from oracle import encoder as oenc
from oracle import network as onet
from oracle import puct as opuct
from oracle import weights as oweights

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # north_star: fp32-accumulate build
BF16_TOL = 2e-2      # north_star: bf16 build, policy probabilities AND value, every network below
ARGMAX_AGREE = 0.99  # north_star: bit-exact argmax move agreement
# The bf16 build = bf16-valued weights x fp16 activation operands (include/caz_b200.h, profiles/r2_precision_ablation.md).
# With bf16 activation operands too (precision="bf16_pure", round 1's build) the value head of the reference's
# random-init net lands at max|dv| 1.6e-2 .. 2.4e-2 depending only on the fp32 summation order: recorded below, not gated
# at north_star's figure.
BF16_PURE_VALUE_TOL = 3e-2


def _planes(n, seed=5):
    fens = oenc.KNOWN_ANSWER_FENS + oenc.synthetic_fens(max(n - 5, 0), seed)
    return np.stack([oenc.canon_input_planes(f) for f in fens[:n]]), fens[:n]


def _make_net(recipe):
    cfg = oweights.keras_config()                                  # 7 x 256, 5x5 stem, 4-filter value conv
    w = oweights.synth_weights(cfg, seed=0, recipe=recipe)
    x, fens = _planes(256)
    want = onet.forward_graph(cfg, w, x)
    return cfg, w, x, fens, want


@pytest.fixture(scope="module")
def net():
    """BN statistics, biases and head scales randomised: folding and bias paths are really exercised."""
    return _make_net("randomized")


@pytest.fixture(scope="module")
def net_default():
    """Config 1: the reference's own random-init ChessModel (Keras default initialisers, identity BN)."""
    return _make_net("keras_default")


@pytest.fixture(scope="module")
def ev_bf16(net):
    ev = caz.Evaluator(net[0], net[1], precision="bf16", max_batch=512)
    yield ev
    ev.close()


@pytest.fixture(scope="module")
def ev_fp32(net):
    ev = caz.Evaluator(net[0], net[1], precision="fp32", max_batch=256)
    yield ev
    ev.close()


# ---- a1: encoder --------------------------------------------------------------------------------
def test_encoder_bit_exact(ev_fp32, golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "encoder_fens.json")))
    want = np.load(os.path.join(golden_dir, "encoder.npz"))["planes_u8"].astype(np.float32)
    got = ev_fp32.encode(caz.encoder.fens_to_records(meta["fens"]))
    assert got.dtype == np.float32 and got.tobytes() == want.tobytes()
    fens = oenc.synthetic_fens(3000, seed=77)                      # > max_batch: exercises chunking
    got = ev_fp32.encode(caz.encoder.fens_to_records(fens))
    want = np.stack([oenc.canon_input_planes(f) for f in fens])
    assert got.tobytes() == want.tobytes()
    assert ev_fp32.encode(np.zeros((0, 68), np.uint8)).shape == (0, 18, 8, 8)


# ---- a9: PUCT ------------------------------------------------------------------------------------
def test_push_priors_and_select_bit_exact(ev_fp32, golden_dir):
    g = np.load(os.path.join(golden_dir, "puct.npz"))
    off = g["offsets"]
    n_nodes = len(off) - 1
    policy = np.zeros((n_nodes, 1968), np.float32)
    for i in range(n_nodes):
        s = slice(off[i], off[i + 1])
        policy[i, g["label_idx"][s]] = g["policy_legal"][s]
    pushed = ev_fp32.push_priors(policy, off, g["label_idx"])
    assert pushed.dtype == np.float64 and np.array_equal(pushed, g["pushed"])
    best = ev_fp32.puct_select(off, g["n"], g["q"], g["p"], g["sum_n"], float(g["c_puct"]), noise=g["noise"],
                               is_root=g["is_root"], noise_eps=float(g["noise_eps"]))
    assert np.array_equal(best, g["best"])
    # full-size: 4096 nodes x up to 218 edges against the oracle loop, including empty nodes
    rng = np.random.RandomState(3)
    sizes = rng.randint(0, 219, size=4096)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    ne = int(off[-1])
    n = rng.randint(0, 50, size=ne).astype(np.int32)
    q = rng.uniform(-1, 1, size=ne)
    p = rng.dirichlet([0.3] * 8, size=ne)[:, 0]
    sum_n = np.asarray([n[off[i]:off[i + 1]].sum() for i in range(4096)], np.int32)
    best = ev_fp32.puct_select(off, n, q, p, sum_n, 1.5)
    for i in range(0, 4096, 7):
        s = slice(off[i], off[i + 1])
        assert best[i] == opuct.select(n[s], q[s], p[s], sum_n[i], 1.5), i
    # the same kernel on device-resident node arrays (caz_puct_select_device: no copies, asynchronous on the handle's stream)
    from chess_alpha_zero_b200 import _lib
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (off, n, q, p, sum_n)]
    d_best = torch.full((4096,), -7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = _lib.load().caz_puct_select_device(ev_fp32._handle, d[0].data_ptr(), 4096, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                            d[4].data_ptr(), None, None, 1.5, 0.0, d_best.data_ptr())
    assert rc == 0
    ev_fp32.sync()
    assert np.array_equal(d_best.cpu().numpy(), best)


# ---- a2/a3: one layer at a time -----------------------------------------------------------------
def _ref_layer(cfg, w, layer, x_nhwc, res, bf16):
    tower, _ = onet.fold_plan(cfg, w)
    _, k, b = tower[layer]
    kt = torch.from_numpy(k.astype(np.float32)).permute(3, 2, 0, 1).contiguous()
    xt = torch.from_numpy(x_nhwc.reshape(-1, 8, 8, x_nhwc.shape[-1])).permute(0, 3, 1, 2).contiguous()
    if bf16:   # the bf16 build's operands: bf16 weights, fp16 activations
        kt, xt = kt.bfloat16().float(), xt.half().float()
    y = F.conv2d(xt.double(), kt.double(), padding=(k.shape[0] - 1) // 2) + torch.from_numpy(b).view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1).reshape(x_nhwc.shape[0], 64, -1)
    if res is not None:
        y = y + torch.from_numpy(res).double()                     # the residual stream is fp32 in both builds
    return torch.relu(y).numpy()


@pytest.mark.parametrize("layer,batch,with_res", [(1, 2, False), (2, 5, True), (0, 3, False), (14, 64, True)])
def test_single_conv_layer_bf16(net, ev_bf16, layer, batch, with_res):
    cfg, w = net[0], net[1]
    rng = np.random.RandomState(layer)
    cin = 18 if layer == 0 else 256
    x = rng.uniform(-1, 1, size=(batch, 64, cin)).astype(np.float32)
    res = rng.uniform(-1, 1, size=(batch, 64, 256)).astype(np.float32) if with_res else None
    got = ev_bf16.debug_conv_layer(layer, x, res)
    want = _ref_layer(cfg, w, layer, x, res, bf16=True)
    # exact-operand reference; what remains is fp32 accumulation order + one bf16 rounding of the output
    assert np.abs(got - want).max() <= 2e-2 * max(1.0, np.abs(want).max()), np.abs(got - want).max()
    assert np.abs(got - want).mean() <= 2e-3


@pytest.mark.parametrize("layer,batch,with_res", [(0, 3, False), (1, 2, False), (2, 3, True)])
def test_single_conv_layer_fp32(net, ev_fp32, layer, batch, with_res):
    cfg, w = net[0], net[1]
    rng = np.random.RandomState(layer)
    cin = 18 if layer == 0 else 256
    x = rng.uniform(-1, 1, size=(batch, 64, cin)).astype(np.float32)
    res = rng.uniform(-1, 1, size=(batch, 64, 256)).astype(np.float32) if with_res else None
    got = ev_fp32.debug_conv_layer(layer, x, res)
    want = _ref_layer(cfg, w, layer, x, res, bf16=False)
    assert np.abs(got - want).max() <= 1e-4


# ---- config 1: whole network, 256 synthetic positions -------------------------------------------
def test_fp32_build_matches_oracle(net, ev_fp32):
    cfg, w, x, fens, (wp, wv) = net
    p, v = ev_fp32.predict_on_batch(x)
    assert p.shape == (256, 1968) and v.shape == (256, 1) and p.dtype == v.dtype == np.float32
    assert np.abs(p - wp).max() <= FP32_TOL and np.abs(v - wv).max() <= FP32_TOL
    assert np.allclose(p.sum(axis=1), 1, atol=1e-5)
    # records in: the encoder runs on the device
    p2, v2 = ev_fp32.predict_records(caz.encoder.fens_to_records(fens))
    assert np.array_equal(p2, p) and np.array_equal(v2, v)


def _bf16_report(tag, p, v, wp, wv, n):
    masks = onet.synthetic_legal_masks(n, seed=1)
    dp, dv = float(np.abs(p - wp).max()), float(np.abs(v - wv).max())
    agree = float((onet.legal_masked_argmax(p, masks) == onet.legal_masked_argmax(wp, masks)).mean())
    raw = float((p.argmax(axis=1) == wp.argmax(axis=1)).mean())
    print(f"{tag}: max|dp|={dp:.3e} max|dv|={dv:.3e} legal-masked argmax agreement {agree:.4f} (unmasked {raw:.4f})")
    return dp, dv, agree


def test_bf16_build_matches_oracle_reference_init(net_default):
    """THE north_star gate for the bf16 build, on the reference's random-init network (config 1)."""
    cfg, w, x, fens, (wp, wv) = net_default
    assert wp.max() > 0.05                                           # the policy is not degenerate
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=256)
    # the bf16-valued weights reach the tensor core exactly, except the ~1e-4 of them below 2^-17 (fp16 subnormals: rounded
    # to a multiple of 2^-24, an absolute error under 3e-8)
    assert ev.inexact_weights < 2e-4 * ev.weight_bytes / 2
    p, v = ev.predict_on_batch(x)
    dp, dv, agree = _bf16_report("bf16 vs fp32 oracle [keras_default]", p, v, wp, wv, len(x))
    assert dp <= BF16_TOL and dv <= BF16_TOL and agree >= ARGMAX_AGREE
    assert np.allclose(p.sum(axis=1), 1, atol=1e-5)
    # informational: the CPU emulation of the same rounding points.  It cannot be a tight gate at network level --
    # a 1e-7 relative perturbation of one layer's accumulators flips bf16 roundings downstream and moves the final
    # activations by as much as the bf16 noise itself (measured); the tight checks are the per-layer tests above.
    ep, evv = onet.forward_bf16(cfg, w, x, fp32_residual=True, fp32_head_input=True, act_format="fp16", w_format="bf16")
    print(f"   vs bf16 emulation: max|dp|={np.abs(p - ep).max():.3e} max|dv|={np.abs(v - evv).max():.3e}")
    assert np.abs(p - ep).max() <= BF16_TOL and np.abs(v - evv).max() <= BF16_TOL
    ev.close()
    ev = caz.Evaluator(cfg, w, precision="fp32", max_batch=256)
    p, v = ev.predict_on_batch(x)
    assert np.abs(p - wp).max() <= FP32_TOL and np.abs(v - wv).max() <= FP32_TOL
    ev.close()


def test_bf16_pure_build_for_the_record(net_default):
    """bf16 weights AND bf16 activation operands: probabilities and argmax inside north_star's bar, the value of this
    net is not (measured 1.6e-2 .. 2.4e-2) -- which is why the bf16 build writes its activation operands as fp16."""
    cfg, w, x, fens, (wp, wv) = net_default
    ev = caz.Evaluator(cfg, w, precision="bf16_pure", max_batch=256)
    p, v = ev.predict_on_batch(x)
    dp, dv, agree = _bf16_report("bf16_pure vs fp32 oracle [keras_default]", p, v, wp, wv, len(x))
    assert dp <= BF16_TOL and dv <= BF16_PURE_VALUE_TOL and agree >= ARGMAX_AGREE
    ps, vs = ev.predict_on_batch(x[:16])                               # small-batch kernel, same operands
    assert np.abs(ps - wp[:16]).max() <= BF16_TOL and np.abs(vs - wv[:16]).max() <= BF16_PURE_VALUE_TOL
    ev.close()


def test_fp16_build_matches_oracle_reference_init(net_default, net):
    """Same tensor-core kernels with fp16 operands (CAZ_PRECISION_FP16): north_star's bf16 gate with a wide margin."""
    for tag, (cfg, w, x, fens, (wp, wv)) in (("keras_default", net_default), ("randomized", net)):
        ev = caz.Evaluator(cfg, w, precision="fp16", max_batch=256)
        p, v = ev.predict_on_batch(x)
        dp, dv, agree = _bf16_report(f"fp16 vs fp32 oracle [{tag}]", p, v, wp, wv, len(x))
        assert dp <= BF16_TOL / 4 and dv <= BF16_TOL / 4 and agree >= ARGMAX_AGREE
        ps, vs = ev.predict_on_batch(x[:16])                          # small-batch kernel, same operands
        assert np.abs(ps - wp[:16]).max() <= BF16_TOL / 4 and np.abs(vs - wv[:16]).max() <= BF16_TOL / 4
        ev.close()


def test_bf16_build_matches_oracle(net, ev_bf16):
    """Stress recipe (random BN statistics, biases, head scales): probabilities and value inside north_star's 2e-2.  Its
    policy rows are near-uniform (max|dp| is 1.4e-3, and only a handful of the 256 rows have a best legal move that leads
    by more than twice that), so: where the oracle's lead exceeds twice the measured error the argmax must ALWAYS agree,
    and over all rows agreement is 96-98 % (profiles/r2_precision_gpu.jsonl; gate 95 %).  On the reference-init net above the bar is
    north_star's 99 % of all rows."""
    cfg, w, x, fens, (wp, wv) = net
    assert wp.max() > 0.01 and wp.std(axis=1).min() > 0              # the policy head is alive
    p, v = ev_bf16.predict_on_batch(x)
    dp, dv, agree = _bf16_report("bf16 vs fp32 oracle [randomized]", p, v, wp, wv, len(x))
    masks = onet.synthetic_legal_masks(len(x), seed=1)
    decisive = onet.decisive_rows(wp, masks, 2 * dp)
    agree_decisive = float((onet.legal_masked_argmax(p, masks) == onet.legal_masked_argmax(wp, masks))[decisive].mean())
    print(f"   {int(decisive.sum())} of {len(x)} rows lead by more than 2 max|dp| = {2 * dp:.1e}: agreement {agree_decisive:.4f}")
    assert dp <= BF16_TOL and dv <= BF16_TOL and agree >= 0.95 and (decisive.sum() == 0 or agree_decisive == 1.0)
    assert np.allclose(p.sum(axis=1), 1, atol=1e-5)
    p2, v2 = ev_bf16.predict_records(caz.encoder.fens_to_records(fens))
    assert np.array_equal(p2, p) and np.array_equal(v2, v)


def test_masked_softmax(net, ev_bf16):
    cfg, w, x, fens, _ = net
    masks = onet.synthetic_legal_masks(64, seed=2)
    p, _ = ev_bf16.predict_on_batch(x[:64])
    pm, _ = ev_bf16.predict_on_batch(x[:64], legal_mask=masks)
    assert np.all(pm[masks == 0] == 0) and np.allclose(pm.sum(axis=1), 1, atol=1e-5)
    want = np.where(masks != 0, p, 0)
    want = want / want.sum(axis=1, keepdims=True)
    assert np.abs(pm - want).max() <= 1e-5


@pytest.mark.parametrize("batch", [1, 2, 3, 16, 17, 48, 100, 148, 149, 255])
def test_ragged_batches_are_consistent(net, ev_bf16, batch):
    """Per-layer kernels (small-batch kernel off): a position's result does not depend on its batch -- nor on whether
    its board group ran as one 256-channel tile or as two 128-channel N tiles (up to 148 boards)."""
    x = net[2]
    ev_bf16.set_small_batch_limit(0)
    try:
        full_p, full_v = ev_bf16.predict_on_batch(x)
        p, v = ev_bf16.predict_on_batch(x[:batch])
    finally:
        ev_bf16.set_small_batch_limit(-1)
    assert np.array_equal(p, full_p[:batch]) and np.array_equal(v, full_v[:batch])


@pytest.mark.parametrize("batch", [1, 5, 73, 100, 148, 149, 255, 296, 297, 300, 512])
def test_fused_tower_matches_per_layer_chain(net, ev_bf16, batch):
    """The one-launch tower (conv_tower.cuh: every CTA pair carries its boards through all layers) against the chain of
    per-layer launches it replaces: bit-identical, with one group per pair and with two groups interleaved."""
    x = np.ascontiguousarray(np.concatenate([net[2], net[2][::-1]])[:batch])
    ev_bf16.set_small_batch_limit(0)
    try:
        ev_bf16.set_fused_tower(0)
        want_p, want_v = ev_bf16.predict_on_batch(x)
        for mode in (-1, 1, 2):
            ev_bf16.set_fused_tower(mode)
            l0 = ev_bf16.kernel_launches
            p, v = ev_bf16.predict_on_batch(x)
            assert ev_bf16.kernel_launches - l0 <= 8, "pack + tower + five head launches"
            assert np.array_equal(p, want_p) and np.array_equal(v, want_v), (batch, mode)
    finally:
        ev_bf16.set_fused_tower(-1)
        ev_bf16.set_small_batch_limit(-1)


def test_fused_tower_soak(net, ev_bf16):
    """Random batch sizes above the small-batch kernel's range, queued back to back for a few seconds with small batches in
    between (the scratch slab and the act_ready barriers are reused launch after launch): every checked result is
    bit-identical to the per-layer chain's."""
    import time
    import torch
    x = np.ascontiguousarray(np.concatenate([net[2], net[2][::-1]]))
    sizes = [73, 100, 148, 149, 256, 296, 297, 300, 512]
    ev_bf16.set_fused_tower(0)
    want = {b: ev_bf16.predict_on_batch(x[:b]) for b in sizes}
    ev_bf16.set_fused_tower(-1)
    d_x = torch.from_numpy(x).cuda()
    d_p = torch.empty((512, 1968), device="cuda")
    d_v = torch.empty((512,), device="cuda")
    torch.cuda.synchronize()
    rng = np.random.default_rng(1)
    t0, launches = time.time(), 0
    while time.time() - t0 < 3.0:
        for _ in range(20):
            b = int(rng.choice(sizes)) if rng.random() < 0.8 else int(rng.integers(1, 73))
            ev_bf16.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
        b = int(rng.choice(sizes))
        ev_bf16.eval_device(d_x.data_ptr(), b, d_p.data_ptr(), d_v.data_ptr())
        ev_bf16.sync()
        launches += 21
        assert np.array_equal(d_p[:b].cpu().numpy(), want[b][0]) and np.array_equal(d_v[:b].cpu().numpy(), want[b][1].reshape(-1)), (b, launches)
    assert launches > 2000


def test_small_batch_persistent_kernel(net_default):
    """Batches <= 64 run the persistent whole-tower kernel (halo tile + shifted UMMA descriptors, weights streamed
    from L2, grid barriers).  Same north_star gate as the per-layer path, and batch-size independent bit for bit."""
    cfg, w, x, fens, (wp, wv) = net_default
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=256)
    ref_p, ref_v = ev.predict_on_batch(x[:64])
    dp, dv, agree = _bf16_report("small-batch kernel vs fp32 oracle [keras_default, 64]", ref_p, ref_v, wp[:64], wv[:64], 64)
    assert dp <= BF16_TOL and dv <= BF16_TOL and agree >= ARGMAX_AGREE
    for batch in (1, 2, 3, 15, 16, 17, 32, 33, 36, 37, 48):
        p, v = ev.predict_on_batch(x[:batch])
        assert np.array_equal(p, ref_p[:batch]) and np.array_equal(v, ref_v[:batch]), batch
    for batch in (65, 72, 73, 74):   # the largest grids: 36 CTA pairs x 2 slices, then 37 two-board tiles x 4 slices = 148 CTAs
        l0 = ev.kernel_launches
        p, v = ev.predict_on_batch(x[:batch])
        assert ev.kernel_launches - l0 == 1, batch
        assert np.array_equal(p[:64], ref_p) and np.array_equal(v[:64], ref_v), batch
        assert np.abs(p - wp[:batch]).max() <= BF16_TOL and np.abs(v - wv[:batch]).max() <= BF16_TOL
    # on request (caz_set_small_batch_limit) up to 148 boards: two-board tiles x two 128-channel slices, the instantiation
    # whose epilogue threads carry 128 channels; same arithmetic, same bits
    ev.set_small_batch_limit(148)
    for batch in (75, 100, 147, 148):
        l0 = ev.kernel_launches
        p, v = ev.predict_on_batch(x[:batch])
        assert ev.kernel_launches - l0 == 1, batch
        assert np.array_equal(p[:64], ref_p) and np.array_equal(v[:64], ref_v), batch
        assert np.abs(p - wp[:batch]).max() <= BF16_TOL and np.abs(v - wv[:batch]).max() <= BF16_TOL
    ev.set_small_batch_limit(-1)
    # against the per-layer kernels: same arithmetic up to fp32 summation order, so bf16-noise level agreement
    ev.set_small_batch_limit(0)
    big_p, big_v = ev.predict_on_batch(x[:64])
    print(f"   small vs per-layer path: max|dp|={np.abs(big_p - ref_p).max():.3e} max|dv|={np.abs(big_v - ref_v).max():.3e}")
    assert np.abs(big_p - ref_p).max() <= BF16_TOL and np.abs(big_v - ref_v).max() <= BF16_TOL
    ev.close()
    # the architecture the reference ships as JSON (3x3 stem) and a 19-block tower through the same kernel
    for variant, n in ((dict(stem_k=3, value_filters=1), 16), (dict(blocks=19), 8)):
        cfg2 = oweights.keras_config(**variant)
        w2 = oweights.synth_weights(cfg2, seed=7, recipe="keras_default")
        want_p, want_v = onet.forward_graph(cfg2, w2, x[:n])
        ev2 = caz.Evaluator(cfg2, w2, precision="bf16", max_batch=32)
        p, v = ev2.predict_on_batch(x[:n])
        assert np.abs(p - want_p).max() <= BF16_TOL and np.abs(v - want_v).max() <= BF16_TOL, variant
        ev2.close()


@pytest.mark.parametrize("precision,tol_p,tol_v", [("bf16", BF16_TOL, BF16_TOL), ("fp16", 4e-3, 4e-3)])
def test_small_batch_kernel_records_masks_and_precisions(net_default, precision, tol_p, tol_v):
    """The persistent kernel's other entrances: board records (its in-kernel encoder builds the stem tile), the
    legal-move mask in its per-tile softmax, fp16 operands."""
    from chess_alpha_zero_b200 import encoder
    cfg, w, x, fens, (wp, wv) = net_default
    ev = caz.Evaluator(cfg, w, precision=precision, max_batch=64)
    for n in (5, 16, 24, 40):
        p, v = ev.predict_on_batch(x[:n])
        assert np.abs(p - wp[:n]).max() <= tol_p and np.abs(v - wv[:n]).max() <= tol_v
        pr, vr = ev.predict_records(encoder.fens_to_records(fens[:n]))
        assert np.array_equal(pr, p) and np.array_equal(vr, v)       # records in == planes in, bit for bit
        masks = onet.synthetic_legal_masks(n, seed=n)
        pm, vm = ev.predict_on_batch(x[:n], legal_mask=masks)
        assert np.all(pm[masks == 0] == 0) and np.allclose(pm.sum(axis=1), 1, atol=1e-5) and np.array_equal(vm, v)
        want = np.where(masks != 0, p, 0)
        assert np.abs(pm - want / want.sum(axis=1, keepdims=True)).max() <= 1e-5
    ev.close()


def test_full_size_batch_properties(net):
    """BASELINE's full size (batch 4096 on one GPU), through size-independent properties: every 256-position block of a
    4096 batch is bit-identical to the 256-position evaluation that is checked against the oracle above; rows are
    probability vectors; values are tanh outputs; a permuted batch gives the permuted result."""
    cfg, w, x, fens, (wp, wv) = net
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=4096)
    base_p, base_v = ev.predict_on_batch(x)
    assert np.abs(base_p - wp).max() <= BF16_TOL
    big = np.ascontiguousarray(np.tile(x, (16, 1, 1, 1)))
    p, v = ev.predict_on_batch(big)
    assert p.shape == (4096, 1968) and v.shape == (4096, 1)
    for blk in range(16):
        assert np.array_equal(p[blk * 256:(blk + 1) * 256], base_p) and np.array_equal(v[blk * 256:(blk + 1) * 256], base_v)
    assert np.allclose(p.sum(axis=1), 1.0, atol=1e-5) and (p >= 0).all() and (np.abs(v) < 1).all()
    perm = np.random.default_rng(0).permutation(4096)
    pp, pv = ev.predict_on_batch(np.ascontiguousarray(big[perm]))
    assert np.array_equal(pp, p[perm]) and np.array_equal(pv, v[perm])
    ev.close()


def test_small_batch_kernel_soak(net_default):
    """Random batch sizes 1..64 back to back for a few seconds (paired and unpaired geometries alternate, flag epochs
    advance): every repeat is bit-identical, and row i does not depend on the batch it came in."""
    import time
    cfg, w, x, fens, _ = net_default
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=64)
    rng = np.random.default_rng(0)
    seen, calls, t0 = {}, 0, time.time()
    while time.time() - t0 < 3.0:
        b, off = int(rng.integers(1, 65)), int(rng.integers(0, 4)) * 64
        p, v = ev.predict_on_batch(x[off:off + b])
        if (b, off) in seen:
            assert np.array_equal(seen[(b, off)][0], p) and np.array_equal(seen[(b, off)][1], v), (b, off, calls)
        else:
            seen[(b, off)] = (p.copy(), v.copy())
        calls += 1
    for off in range(0, 256, 64):
        full = ev.predict_on_batch(x[off:off + 64])
        for (b, o), (p, v) in seen.items():
            if o == off:
                assert np.array_equal(p, full[0][:b]) and np.array_equal(v, full[1][:b]), (b, off)
    assert calls > 1000
    ev.close()


def test_empty_batch_and_chunking(net, ev_bf16):
    p, v = ev_bf16.predict_on_batch(np.zeros((0, 18, 8, 8), np.float32))
    assert p.shape == (0, 1968) and v.shape == (0, 1)
    x = np.concatenate([net[2]] * 5)[:1100]                        # > max_batch (512): 3 chunks
    p, v = ev_bf16.predict_on_batch(x)
    base_p, base_v = ev_bf16.predict_on_batch(net[2])
    assert np.array_equal(p[256:512], base_p) and np.array_equal(p[1024:1100], base_p[:76])


def test_submit_collect_pipelined_api(net, ev_bf16):
    """Two batches in flight through caz_submit / caz_collect give exactly what the synchronous call gives."""
    from chess_alpha_zero_b200 import _lib
    x = net[2]
    want = [ev_bf16.predict_on_batch(x[:200]), ev_bf16.predict_on_batch(x[56:256])]
    pins = [(_lib.PinnedArray((200, 18, 8, 8), np.float32), _lib.PinnedArray((200, 1968), np.float32),
             _lib.PinnedArray((200, 1), np.float32)) for _ in range(2)]
    pins[0][0].array[:] = x[:200]
    pins[1][0].array[:] = x[56:256]
    base = ev_bf16.submit(pins[0][0].array, pins[0][1].array, pins[0][2].array)     # tickets count up per handle
    ev_bf16.collect(base)
    with pytest.raises(caz.CazError):
        ev_bf16.collect(base)                                                          # a ticket is collected once
    base += 1
    for rounds in range(3):
        t = [ev_bf16.submit(pins[i][0].array, pins[i][1].array, pins[i][2].array) for i in range(2)]
        assert t[1] == t[0] + 1 and t[0] == 2 * rounds + base
        for i in range(2):
            ev_bf16.collect(t[i])
            assert np.array_equal(pins[i][1].array, want[i][0]) and np.array_equal(pins[i][2].array, want[i][1])
        for i in range(2):
            pins[i][1].array[:] = 0
    with pytest.raises(caz.CazError):
        ev_bf16.submit(np.zeros((513, 18, 8, 8), np.float32), np.zeros((513, 1968), np.float32), np.zeros((513, 1), np.float32))
    # four submissions in flight (what the self-play driver does with four game groups): each ticket completes on its own
    # event, in any collection order; the 17th outstanding submission is refused
    more = [(_lib.PinnedArray((200, 1968), np.float32), _lib.PinnedArray((200, 1), np.float32)) for _ in range(4)]
    t = [ev_bf16.submit(pins[i & 1][0].array, more[i][0].array, more[i][1].array) for i in range(4)]
    for i in (2, 0, 3, 1):
        ev_bf16.collect(t[i])
        assert np.array_equal(more[i][0].array, want[i & 1][0]) and np.array_equal(more[i][1].array, want[i & 1][1])
    t = [ev_bf16.submit(pins[0][0].array, more[0][0].array, more[0][1].array) for i in range(16)]
    with pytest.raises(caz.CazError):
        ev_bf16.submit(pins[0][0].array, more[0][0].array, more[0][1].array)
    for tk in t:
        ev_bf16.collect(tk)


def test_submit_records_pipelined_api(net, ev_bf16):
    """caz_submit_records: two batches of board records in flight (what the self-play driver's two game groups do) give
    exactly what the synchronous record call gives, which in turn is the plane call's result."""
    from chess_alpha_zero_b200 import _lib, encoder
    fens = net[3]
    recs = encoder.fens_to_records(fens)
    want = [ev_bf16.predict_records(recs[:200]), ev_bf16.predict_records(recs[56:256])]
    base_p, base_v = ev_bf16.predict_on_batch(net[2])
    assert np.array_equal(want[0][0], base_p[:200]) and np.array_equal(want[1][1], base_v[56:256])
    pins = [(_lib.PinnedArray((200, 68), np.uint8), _lib.PinnedArray((200, 1968), np.float32),
             _lib.PinnedArray((200,), np.float32)) for _ in range(2)]
    pins[0][0].array[:] = recs[:200]
    pins[1][0].array[:] = recs[56:256]
    base = ev_bf16.submit_records(pins[0][0].array, pins[0][1].array, pins[0][2].array) + 1
    ev_bf16.collect(base - 1)
    for rounds in range(3):
        t = [ev_bf16.submit_records(pins[i][0].array, pins[i][1].array, pins[i][2].array) for i in range(2)]
        assert t[1] == t[0] + 1 and t[0] == 2 * rounds + base
        for i in range(2):
            ev_bf16.collect(t[i])
            assert np.array_equal(pins[i][1].array, want[i][0]) and np.array_equal(pins[i][2].array, want[i][1].reshape(-1))
        for i in range(2):
            pins[i][1].array[:] = 0


def test_shipped_json_variant_and_reload(net):
    cfg = oweights.keras_config(stem_k=3, value_filters=1)         # what data/model/model_best_config.json holds
    w = oweights.synth_weights(cfg, seed=4, recipe="randomized")
    x = net[2][:32]
    wp, wv = onet.forward_graph(cfg, w, x)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=64)
    p, v = ev.predict_on_batch(x)
    assert np.abs(p - wp).max() <= BF16_TOL and np.abs(v - wv).max() <= BF16_TOL
    w2 = oweights.synth_weights(cfg, seed=5, recipe="randomized")  # hot reload, handle stays open
    ev.reload(w2)
    wp2, wv2 = onet.forward_graph(cfg, w2, x)
    p2, v2 = ev.predict_on_batch(x)
    assert np.abs(p2 - wp2).max() <= BF16_TOL and np.abs(v2 - wv2).max() <= BF16_TOL
    assert np.abs(p2 - p).max() > 1e-4
    ev.close()


def test_keras_default_init_19_blocks(net):
    cfg = oweights.keras_config(blocks=19)                          # config 5 architecture
    w = oweights.synth_weights(cfg, seed=6, recipe="keras_default")
    x = net[2][:16]
    wp, wv = onet.forward_graph(cfg, w, x)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=16)
    p, v = ev.predict_on_batch(x)
    assert np.abs(p - wp).max() <= BF16_TOL and np.abs(v - wv).max() <= BF16_TOL
    ev.close()


def test_per_layer_path_19_blocks_256_boards(net):
    """Config 5's network (19 x 256) through the PER-LAYER kernels (CTA-pair implicit GEMM, PDL chain, heads as
    tensor-core GEMMs): 256 distinct boards against the oracle, and the small-batch kernel's answer for the first 64."""
    cfg = oweights.keras_config(blocks=19)
    w = oweights.synth_weights(cfg, seed=6, recipe="keras_default")
    x = net[2]
    wp, wv = onet.forward_graph(cfg, w, x)
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=256)
    ev.set_fused_tower(0)                                              # one launch per convolution
    l0 = ev.kernel_launches
    p, v = ev.predict_on_batch(x)
    assert ev.kernel_launches - l0 >= 1 + 39, "256 boards must take the per-layer path (one launch per convolution)"
    dp, dv, agree = _bf16_report("per-layer path, 19x256, 256 boards", p, v, wp, wv, len(x))
    assert dp <= BF16_TOL and dv <= BF16_TOL and agree >= ARGMAX_AGREE
    assert np.allclose(p.sum(axis=1), 1, atol=1e-5)
    ev.set_fused_tower(-1)                                             # the default: the whole tower in one launch
    l0 = ev.kernel_launches
    pf, vf = ev.predict_on_batch(x)
    assert ev.kernel_launches - l0 <= 8
    assert np.array_equal(pf, p) and np.array_equal(vf, v), "fused tower and per-layer chain must agree bit for bit"
    ps, vs = ev.predict_on_batch(x[:64])                               # persistent kernel: same operands, other order
    assert np.abs(ps - p[:64]).max() <= BF16_TOL and np.abs(vs - v[:64]).max() <= BF16_TOL
    ev.set_small_batch_limit(0)
    p2, v2 = ev.predict_on_batch(x[:150])                              # per-layer again, ragged: bit-identical rows
    assert np.array_equal(p2, p[:150]) and np.array_equal(v2, v[:150])
    ev.close()


def test_4096_distinct_positions_against_oracle(net_default):
    """BASELINE's full batch with 4096 DIFFERENT positions: a random 512-row subset is checked against the oracle, every
    row is a probability vector, and the result does not depend on where in the batch a position sits."""
    cfg, w = net_default[0], net_default[1]
    fens = oenc.synthetic_fens(4096, seed=2024)
    assert len(set(fens)) >= 4090
    x = np.stack([oenc.canon_input_planes(f) for f in fens])
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=4096)
    p, v = ev.predict_on_batch(x)
    assert p.shape == (4096, 1968) and np.allclose(p.sum(axis=1), 1.0, atol=1e-5) and (np.abs(v) < 1).all()
    rows = np.sort(np.random.default_rng(5).choice(4096, size=512, replace=False))
    wp, wv = onet.forward_graph(cfg, w, x[rows])
    dp, dv, agree = _bf16_report("4096 distinct positions, 512-row subset", p[rows], v[rows], wp, wv, 512)
    assert dp <= BF16_TOL and dv <= BF16_TOL and agree >= ARGMAX_AGREE
    pr, vr = ev.predict_records(caz.encoder.fens_to_records(fens))    # device-side encoder at full size
    assert np.array_equal(pr, p) and np.array_equal(vr, v)
    p2, v2 = ev.predict_on_batch(x[rows])                              # the same boards in a 512 batch
    assert np.array_equal(p2, p[rows]) and np.array_equal(v2, v[rows])
    ev.close()


# ---- a7/a8: the pipe protocol, ours and the reference's own server class ------------------------
_REFERENCE_API_KEEPALIVE = []


def _reference_api_class():
    """The reference's UNMODIFIED ChessModelAPI: from the checkout in the build container, or from the copy
    __graft_entry__.build() stages under baseline/_ref/ (git-ignored; it travels to the GPU box with the snapshot)."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for src in ("/root/reference/src", os.path.join(root, "baseline", "_ref")):
        if os.path.exists(os.path.join(src, "chess_zero", "agent", "api_chess.py")):
            if src not in sys.path:
                sys.path.insert(0, src)
            return importlib.import_module("chess_zero.agent.api_chess").ChessModelAPI, src
    return None, None


def test_reference_chessmodelapi_runs_on_the_cuda_evaluator(net):
    """Drop-in proof on hardware: the reference's own prediction_worker thread (api_chess.py:51-69) calls
    self.agent_model.model.predict_on_batch on OUR ChessModel; 16 client threads speak its pipe protocol."""
    RefAPI, src = _reference_api_class()
    if RefAPI is None:
        pytest.skip("no copy of the reference's api_chess.py (run __graft_entry__.build() where /root/reference exists)")
    assert "chess-alpha-zero_b200" not in (RefAPI.__module__ or "") and RefAPI is not caz.ChessModelAPI
    cfg, w, x, fens, (wp, wv) = net
    model = caz.ChessModel(caz.Config("mini"), precision="bf16", max_batch=64)
    model._install(cfg, w)
    api = RefAPI(model)                                                # reference code from here on
    pipes = [api.create_pipe() for _ in range(16)]
    api.start()
    results = {}

    def client(i, pipe):
        out = []
        for j in range(i, 128, 16):
            pipe.send(x[j])
            p, v = pipe.recv()
            out.append((j, p, v))
        results[i] = out

    ts = [threading.Thread(target=client, args=(i, p)) for i, p in enumerate(pipes)]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    assert len(results) == 16 and sum(len(o) for o in results.values()) == 128
    for out in results.values():
        for j, p, v in out:
            assert isinstance(v, float) and p.shape == (1968,) and p.dtype == np.float32
            assert np.abs(p - wp[j]).max() <= BF16_TOL and abs(v - float(wv[j, 0])) <= BF16_TOL
    print(f"   reference ChessModelAPI from {src}: 128 requests answered by the CUDA evaluator")
    model.model.close()
    # the reference's worker thread has no way to stop: keep its pipes open so that it sleeps in connection.wait() for the
    # rest of the session instead of dying with EOFError inside some later test
    _REFERENCE_API_KEEPALIVE.append((api, pipes))


@pytest.mark.parametrize("transport", ["shm"])
def test_shm_transport_on_the_real_evaluator(net, transport):
    """create_pipe() clients over the shared-memory transport (shm_pipes.py), real evaluator behind the server."""
    cfg, w, x, fens, (wp, wv) = net
    model = caz.ChessModel(caz.Config("mini"), precision="bf16", max_batch=64, transport=transport)
    model._install(cfg, w)
    pipes = model.get_pipes(24)
    results = {}

    def client(i, pipe):
        out = []
        for j in range(i, 240, 24):
            pipe.send(x[j])
            p, v = pipe.recv()
            out.append((j, np.array(p, copy=True), v))
        results[i] = out

    ts = [threading.Thread(target=client, args=(i, p)) for i, p in enumerate(pipes)]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    assert len(results) == 24 and model.api.positions == 240
    for out in results.values():
        for j, p, v in out:
            assert isinstance(v, float) and p.shape == (1968,) and p.dtype == np.float32
            assert np.abs(p - wp[j]).max() <= BF16_TOL and abs(v - float(wv[j, 0])) <= BF16_TOL
    model.api.stop()
    model.model.close()


def test_chessmodel_pipes_end_to_end(net, tmp_path):
    cfg, w, x, fens, (wp, wv) = net
    cpath, wpath = str(tmp_path / "model_best_config.json"), str(tmp_path / "model_best_weight.npz")
    json.dump(cfg, open(cpath, "wt"))
    keras_graph.save_weight_file(wpath, w)
    model = caz.ChessModel(caz.Config("mini"), precision="bf16", max_batch=64)
    assert model.load(cpath, wpath) and model.digest == caz.ChessModel.fetch_digest(wpath)
    pipes = model.get_pipes(16)
    results = {}

    def client(i, pipe):
        out = []
        for j in range(i, 64, 16):
            pipe.send(x[j])
            p, v = pipe.recv()
            out.append((j, p, v))
        results[i] = out

    ts = [threading.Thread(target=client, args=(i, p)) for i, p in enumerate(pipes)]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    assert len(results) == 16
    for out in results.values():
        for j, p, v in out:
            assert isinstance(v, float) and p.shape == (1968,) and p.dtype == np.float32
            assert np.abs(p - wp[j]).max() <= BF16_TOL and abs(v - float(wv[j, 0])) <= BF16_TOL
    assert model.api.positions == 64
    model.api.stop()
    model.model.close()
