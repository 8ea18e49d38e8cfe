"""The host-side batched search (include/caz_mcts.h) driving the CUDA evaluator through caz_eval_records: what the
leaves get back is what the oracle computes for the same positions, and the games advance.  Needs a B200 (-m gpu)."""
import numpy as np
import pytest

import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import labels, mcts
from oracle import encoder as oenc
from oracle import network as onet
from oracle import weights as oweights

pytestmark = pytest.mark.gpu


def test_selfplay_rounds_through_the_evaluator():
    cfg = oweights.keras_config()
    w = oweights.synth_weights(cfg, seed=0, recipe="randomized")
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=8 * 16)
    mc = mcts.MctsConfig.normal(simulation_num_per_move=64, search_threads=16, continuous=0, max_game_length=6)
    s = caz.SelfPlaySearch(8, mc, seed=4, n_threads=2)
    checked = 0
    for rnd in range(400):
        rec = s.collect()
        if len(rec) == 0:
            break
        policy, value = ev.predict_records(rec)
        assert policy.shape == (len(rec), labels.N_LABELS) and np.allclose(policy.sum(1), 1.0, atol=1e-4)
        if rnd in (0, 3):
            # the records the search hands out encode to the reference's planes, and the device result for them is the
            # oracle's (bf16 gate of north_star)
            planes = ev.encode(rec)
            fens = [s.fen(g) for g in range(8)] if rnd == 0 else None
            if fens:
                want_planes = np.stack([oenc.canon_input_planes(f) for f in fens])
                assert planes.tobytes() == want_planes.tobytes()
            wp, wv = onet.forward_graph(cfg, w, planes)
            assert np.abs(policy - wp).max() <= 2e-2 and np.abs(value - wv).max() <= 2e-2
            checked += 1
        s.apply(policy, value)
    st = s.stats()
    assert checked == 2 and st["games_finished"] == 8 and st["moves_played"] >= 8 * 6 - 8
    assert st["descents"] == st["leaves_evaluated"] + st["terminal_hits"]
    s.close()
    ev.close()


def test_device_pushed_priors_are_bit_identical_and_build_the_same_trees():
    """caz_submit_records_priors (legal-move labels in, one float64 prior per legal move out) against the full-policy path:
    the priors equal caz_push_priors of the policy rows bit for bit, and two searches fed either way stay identical."""
    from chess_alpha_zero_b200 import _lib
    cfg = oweights.keras_config()
    w = oweights.synth_weights(cfg, seed=0, recipe="randomized")
    ev = caz.Evaluator(cfg, w, precision="bf16", max_batch=6 * 16)
    mc = mcts.MctsConfig.normal(simulation_num_per_move=48, search_threads=16, continuous=0, max_game_length=4)
    a = caz.SelfPlaySearch(6, mc, seed=9, n_threads=2)
    b = caz.SelfPlaySearch(6, mc, seed=9, n_threads=2)
    cap = b.capacity
    off, lab = _lib.PinnedArray((cap + 1,), np.int32), _lib.PinnedArray((cap * 256,), np.int32)
    pri, val = _lib.PinnedArray((cap * 256,), np.float64), _lib.PinnedArray((cap,), np.float32)
    rounds = 0
    for rnd in range(300):
        rec_a = a.collect()
        rec_b, o, l = b.collect_csr(off.array, lab.array)
        assert np.array_equal(rec_a, rec_b)
        if len(rec_a) == 0:
            break
        policy, value = ev.predict_records(rec_a)
        ev.collect(ev.submit_records_priors(rec_b, off.array, lab.array, pri.array, val.array))
        n_edges = int(o[-1])
        want = ev.push_priors(policy, o, np.ascontiguousarray(np.maximum(l, 0)))
        assert (l >= 0).all() and np.array_equal(pri.array[:n_edges], want)
        assert np.array_equal(val.array[:len(rec_b)], value.reshape(-1))
        a.apply(policy, value)
        b.apply_priors(pri.array, val.array)
        rounds += 1
    assert rounds > 10 and a.stats() == b.stats() and a.stats()["games_finished"] == 6
    assert [a.fen(g) for g in range(6)] == [b.fen(g) for g in range(6)]
    a.close()
    b.close()
    ev.close()
