"""CPU checks of the host-side batched search (include/caz_mcts.h): the chess rules underneath (published perft
counts, python-chess's move order / FEN / draw-claim behaviour restated by hand -- python-chess itself is not
installed here, so these are known answers, not a live comparison) and the search's own invariants."""
import os
import re

import numpy as np
import pytest

import chess_alpha_zero_b200 as caz
from chess_alpha_zero_b200 import encoder, labels, mcts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "caz_mcts.h")).read()
    declared = set(re.findall(r"\b(caz_(?:mcts|chess)_\w+)\s*\(", header))
    assert declared == set(mcts.SYMBOLS)
    lib = mcts.load()
    for name in declared:
        assert hasattr(lib, name), name


# https://www.chessprogramming.org/Perft_Results -- the standard move-generator known answers
PERFT = [
    (START, [20, 400, 8902, 197281]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", [48, 2039, 97862]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890]),
]


@pytest.mark.parametrize("fen,counts", PERFT)
def test_perft_known_answers(fen, counts):
    for depth, want in enumerate(counts, start=1):
        assert mcts.perft(fen, depth) == want, (fen, depth)


def test_legal_move_order_follows_python_chess():
    # list(chess.Board().legal_moves) of python-chess: knights from the highest square, then pawn pushes h..a, doubles h..a
    assert mcts.legal_moves(START) == ("g1h3 g1f3 b1c3 b1a3 h2h3 g2g3 f2f3 e2e3 d2d3 c2c3 b2b3 a2a3 "
                                       "h2h4 g2g4 f2f4 e2e4 d2d4 c2c4 b2b4 a2a4").split()
    # pieces from the highest square down, targets from the highest square down, castling last (h-side first)
    got = mcts.legal_moves("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    want = ["h1h8", "h1h7", "h1h6", "h1h5", "h1h4", "h1h3", "h1h2", "h1g1", "h1f1",
            "e1f2", "e1e2", "e1d2", "e1f1", "e1d1",
            "a1a8", "a1a7", "a1a6", "a1a5", "a1a4", "a1a3", "a1a2", "a1d1", "a1c1", "a1b1", "e1g1", "e1c1"]
    assert got == want
    # in check: king moves first, then captures / blocks of the single checker; promotions q, r, b, n
    got = mcts.legal_moves("4k3/8/8/8/8/8/4r3/R3K3 w Q - 0 1")
    assert got == ["e1e2", "e1f1", "e1d1"]
    assert mcts.legal_moves("8/P6k/8/8/8/8/8/K7 w - - 0 1")[-4:] == ["a7a8q", "a7a8r", "a7a8b", "a7a8n"]


def test_fen_writes_en_passant_only_when_a_capture_is_legal():
    # python-chess: board.fen() defaults to en_passant="legal"
    assert mcts.play(START, ["e2e4"])[1] == "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq - 0 1"
    assert mcts.play(START, "e2e4 d7d5 e4e5 f7f5".split())[1] == "rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3"
    # the capturing pawn is pinned against its king: not legal, not written
    fen = mcts.play("8/8/8/8/k2p3R/8/4P3/4K3 w - - 0 1", ["e2e4"])[1]
    assert fen.split()[3] == "-"
    # castling rights disappear with king / rook moves and rook captures
    assert mcts.play("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", ["a1a8"])[1].split()[2] == "Kk"
    assert mcts.play("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", ["e1g1"])[1].startswith("r3k2r/8/8/8/8/8/8/R4RK1 b kq")


def test_result_with_claimed_draws():
    assert mcts.play(START, "f2f3 e7e5 g2g4 d8h4".split())[0] == "0-1"                     # fool's mate
    assert mcts.play("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")[0] == "1/2-1/2"                     # stalemate
    assert mcts.play("8/8/4k3/8/8/3K4/8/8 w - - 0 1")[0] == "1/2-1/2"                      # K v K
    assert mcts.play("8/8/4k3/8/8/3KN3/8/8 w - - 0 1")[0] == "1/2-1/2"                     # KN v K
    assert mcts.play("8/8/4k3/8/8/3KNN2/8/8 w - - 0 1")[0] == "*"                          # KNN v K is not claimed
    assert mcts.play("8/8/2b1k3/8/8/3KB3/8/8 w - - 0 1")[0] == "*"                         # opposite-coloured bishops
    assert mcts.play("8/8/3bk3/8/8/3KB3/8/8 w - - 0 1")[0] == "1/2-1/2"                    # same-coloured bishops
    assert mcts.play("8/8/4k3/8/8/3KR3/8/8 w - - 99 80")[0] == "*"
    assert mcts.play("8/8/4k3/8/8/3KR3/8/8 w - - 100 80")[0] == "1/2-1/2"                  # fifty-move claim
    shuffle = "g1f3 g8f6 f3g1 f6g8".split()
    assert mcts.play(START, shuffle + shuffle[:2])[0] == "*"
    # after 7 plies ...Ng8 would bring the start position for the third time: claimable already (python-chess
    # can_claim_threefold_repetition looks one move ahead)
    assert mcts.play(START, shuffle + shuffle[:3])[0] == "1/2-1/2"
    assert mcts.play(START, shuffle + shuffle)[0] == "1/2-1/2"
    # an irreversible move starts the count again (the move stack keeps occurrences per reversible stretch)
    after_pawns = shuffle + ["e2e4", "e7e5"]
    assert mcts.play(START, after_pawns + shuffle)[0] == "*"
    assert mcts.play(START, after_pawns + shuffle + shuffle[:2])[0] == "*"
    assert mcts.play(START, after_pawns + shuffle + shuffle[:3])[0] == "1/2-1/2"
    assert mcts.play(START, after_pawns + shuffle + "d2d4 g8f6 g1f3 f6g8 f3g1".split())[0] == "*"
    with pytest.raises(ValueError):
        mcts.play(START, ["e2e5"])


def test_records_agree_with_the_python_fen_path():
    fens = [START]
    line = "e2e4 d7d5 e4e5 f7f5 e5f6 g8f6 g1f3 e7e6 f1b5 c7c6 e1g1 f8d6 b5a4 e8g8 d2d4 b8d7 c1g5 h7h6".split()
    for i in range(1, len(line) + 1):
        fens.append(mcts.play(START, line[:i])[1])
    want = encoder.fens_to_records(fens)
    got = np.stack([mcts.record(f) for f in fens])
    assert np.array_equal(got, want)


def _fake_net(records, n_labels=labels.N_LABELS):
    """A deterministic position-dependent 'network': policy and value from a hash of the board record."""
    h = (records.astype(np.uint64) * np.arange(1, 69, dtype=np.uint64)).sum(1)
    pol = np.empty((len(records), n_labels), np.float32)
    base = np.arange(n_labels, dtype=np.uint64)
    for i, hv in enumerate(h):
        x = ((base * np.uint64(2654435761) + hv * np.uint64(40503)) % np.uint64(1009)).astype(np.float32) + 1.0
        pol[i] = x / x.sum()
    val = ((h % np.uint64(2001)).astype(np.float32) - 1000.0) / 5000.0
    return pol, val


def _run(search, rounds):
    for _ in range(rounds):
        rec = search.collect()
        if len(rec) == 0:
            break
        search.apply(*_fake_net(rec))


def test_search_is_deterministic_and_independent_of_thread_count():
    out = []
    for threads in (1, 3):
        cfg = mcts.MctsConfig.normal(simulation_num_per_move=40, search_threads=8)
        s = caz.SelfPlaySearch(6, cfg, seed=11, n_threads=threads)
        _run(s, 60)
        out.append((s.stats(), [s.fen(g) for g in range(6)]))
        s.close()
    assert out[0] == out[1]
    assert out[0][0]["moves_played"] > 0 and out[0][0]["descents"] == out[0][0]["leaves_evaluated"] + out[0][0]["terminal_hits"]


def test_root_statistics_and_prior_normalisation():
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=200, search_threads=16, noise_eps=0.0)
    s = caz.SelfPlaySearch(1, cfg, seed=5, n_threads=1)
    assert s.root_edges(0) is None
    rec = s.collect()
    assert rec.shape == (1, 68)                      # the first round of a move is the root alone
    pol, val = _fake_net(rec)
    s.apply(pol, val)
    lab, n, q, p = s.root_edges(0)
    moves = mcts.legal_moves(START)
    assert [labels.LABELS[i] for i in lab] == moves  # edges in legal-move order, white: labels as they are
    # select_action_q_and_u's prior push (player_chess.py:267-275): p = P[label] / (1e-8 + sum), float64
    raw = pol[0][lab].astype(np.float64)
    tot = 1e-8
    for x in raw:
        tot += x
    assert np.array_equal(p, raw / tot) and n.sum() == 0
    rec = s.collect()
    assert 1 <= len(rec) <= 16                       # up to search_threads descents, virtual loss spreads them
    s.apply(*_fake_net(rec))
    lab, n, q, p = s.root_edges(0)
    assert n.sum() == len(rec) and (n >= 0).all()
    # with virtual loss 3 and equal-ish priors the 16 descents of one round must not all pile onto one move
    assert (n > 0).sum() > 1
    s.close()


def test_black_to_move_uses_the_flipped_label():
    """For black the network's (canonical) policy is permuted with Config.unflipped_index (player_chess.py:232-233):
    mass on canonical e2e4 must land on black's e7e5."""
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=10, noise_eps=0.0)
    s = caz.SelfPlaySearch(1, cfg, seed=1, n_threads=1)
    s.set_position(0, "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq - 0 1")
    rec = s.collect()
    pol = np.full((1, labels.N_LABELS), 1e-6, np.float32)
    pol[0, labels.LABEL_INDEX["e2e4"]] = 1.0
    s.apply(pol, np.zeros(1, np.float32))
    lab, n, q, p = s.root_edges(0)
    best = labels.LABELS[lab[int(np.argmax(p))]]
    assert best == "e7e5" and p.max() > 0.99
    s.close()


def test_search_finds_mate_in_one_and_ends_the_game():
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=300, noise_eps=0.0, tau_decay_rate=0.01, continuous=0,
                                 resign_enabled=0)
    s = caz.SelfPlaySearch(1, cfg, seed=2, n_threads=1)
    s.set_position(0, "6k1/5ppp/8/8/8/8/8/R5K1 w - - 0 1")
    uniform = lambda k: (np.full((k, labels.N_LABELS), 1.0 / labels.N_LABELS, np.float32), np.zeros(k, np.float32))  # noqa: E731
    for _ in range(200):
        rec = s.collect()
        if len(rec) == 0:
            break
        s.apply(*uniform(len(rec)))
    assert s.result(0) == "1-0" and s.fen(0).startswith("R5k1/")
    st = s.stats()
    assert st["games_finished"] == 1 and st["white_wins"] == 1 and st["terminal_hits"] > 0
    s.close()


def test_games_run_to_completion():
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=12, search_threads=4, max_game_length=24, continuous=0)
    s = caz.SelfPlaySearch(5, cfg, seed=9, n_threads=2)
    for _ in range(2000):
        rec = s.collect()
        if len(rec) == 0:
            break
        s.apply(*_fake_net(rec))
    st = s.stats()
    assert st["games_finished"] == 5 == st["white_wins"] + st["black_wins"] + st["draws"]
    assert all(s.result(g) != "*" for g in range(5))
    with pytest.raises(ValueError):
        s.apply(np.zeros((3, labels.N_LABELS), np.float32), np.zeros(3, np.float32))
    s.close()


def test_play_data_rows_match_the_reference_format():
    """self_play_buffer's data (self_play.py:135-152): one [fen, policy, z] row per move, z from the mover's side."""
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=16, search_threads=4, max_game_length=12, continuous=0,
                                 record_play_data=1, resign_enabled=0)
    s = caz.SelfPlaySearch(3, cfg, seed=21, n_threads=1)
    for _ in range(1000):
        rec = s.collect()
        if len(rec) == 0:
            break
        s.apply(*_fake_net(rec))
    rows = s.drain_play_data()
    st = s.stats()
    assert len(rows) == st["moves_played"] == 3 * 12 and s.drain_play_data() == []
    results = {"1-0": 1, "0-1": -1, "1/2-1/2": 0}
    game_of_row = 0
    for k, (fen, policy, z) in enumerate(rows):
        if k and fen.startswith("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w"):
            game_of_row += 1
        assert len(policy) == labels.N_LABELS and abs(sum(policy) - 1.0) < 1e-5
        legal = {labels.LABEL_INDEX[m] for m in mcts.legal_moves(fen)}
        assert {i for i, p in enumerate(policy) if p > 0} <= legal
        white_to_move = fen.split()[1] == "w"
        assert z in (-1, 0, 1)
    # z is consistent inside a game: white's rows carry the white result, black's the negation
    assert game_of_row == 2
    per_game = [rows[i * 12:(i + 1) * 12] for i in range(3)]
    for g, chunk in enumerate(per_game):
        want_white = results[s.result(g)] if s.result(g) in results else None
        for fen, _, z in chunk:
            assert z == (want_white if fen.split()[1] == "w" else -want_white)
    s.close()


def test_root_noise_is_dirichlet():
    """np.random.dirichlet([0.3] * n) (player_chess.py:285-286) from the search's own fast generator: moments of
    Dirichlet(alpha, ..., alpha): mean 1/n, variance (n - 1) / (n^2 (n alpha + 1)); marginals are Beta(alpha, (n-1) alpha)."""
    n, alpha = 30, 0.3
    x = mcts.dirichlet_draws(7, alpha, n, 200000)
    assert np.allclose(x.sum(1), 1.0, atol=1e-12) and (x >= 0).all()
    assert np.abs(x.mean(0) - 1.0 / n).max() < 6e-4
    want_var = (n - 1) / (n * n * (n * alpha + 1))
    assert np.abs(x.var(0) / want_var - 1).max() < 0.03
    # the spiky shape that matters for exploration: P(x_i > 0.2) for Beta(0.3, 8.7)
    from scipy import stats
    assert abs((x[:, 0] > 0.2).mean() - stats.beta.sf(0.2, alpha, (n - 1) * alpha)) < 2e-3
    assert not np.array_equal(mcts.dirichlet_draws(8, alpha, n, 4), x[:4])


def test_double_collect_is_refused_and_the_search_goes_on():
    """collect() twice without an apply() (a driver retrying after a failed evaluator call) must not spin: status 1,
    nothing changes, and the original batch can still be applied."""
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=8, search_threads=4)
    s = caz.SelfPlaySearch(2, cfg, seed=3, n_threads=4)
    rec = s.collect().copy()
    assert len(rec) == 2
    before = s.stats()
    with pytest.raises(RuntimeError):
        s.collect()
    assert s.stats() == before
    s.apply(*_fake_net(rec))
    _run(s, 30)
    assert s.stats()["moves_played"] > 0
    s.close()


def test_one_simulation_per_move_is_rejected():
    """With a single simulation only the root is expanded: there are no visit counts to choose a move from (the
    reference would raise in np.random.choice)."""
    with pytest.raises((RuntimeError, ValueError)):
        caz.SelfPlaySearch(1, mcts.MctsConfig.normal(simulation_num_per_move=1), seed=1, n_threads=1)


def test_adjudication_overrides_a_result_at_the_length_limit():
    """self_play_buffer adjudicates by material whenever num_halfmoves >= max_game_length, even when the step itself
    ended the game (self_play.py:127-133): a mate delivered on the last allowed ply is scored by material count."""
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=300, noise_eps=0.0, tau_decay_rate=0.01, continuous=0,
                                 resign_enabled=0, max_game_length=1)
    s = caz.SelfPlaySearch(1, cfg, seed=2, n_threads=1)
    s.set_position(0, "6k1/5ppp/8/8/8/8/8/R5K1 w - - 0 1")      # Ra8 mates, but black has 3 pawns for the rook: 5 vs 3
    uniform = lambda k: (np.full((k, labels.N_LABELS), 1.0 / labels.N_LABELS, np.float32), np.zeros(k, np.float32))  # noqa: E731
    for _ in range(200):
        rec = s.collect()
        if len(rec) == 0:
            break
        s.apply(*uniform(len(rec)))
    st = s.stats()
    assert st["games_finished"] == 1 and st["adjudications"] == 1 and s.fen(0).startswith("R5k1/")
    assert s.result(0) == "1-0"                                   # tanh(3 * (5 + 3 - 3 - 3) / (5 + 3 + 3 + 3)) > 0.01
    s.close()


def test_apply_priors_builds_the_same_trees_as_apply():
    """The CSR path (collect_csr -> per-edge priors pushed elsewhere -> apply_priors) against apply(): the prior push is
    restated here the way push_priors_kernel does it (float32 -> float64, sequential sum from 1e-8, labels < 0 as 0)."""
    cfg = mcts.MctsConfig.normal(simulation_num_per_move=30, search_threads=8)
    a = caz.SelfPlaySearch(4, cfg, seed=21, n_threads=2)
    b = caz.SelfPlaySearch(4, cfg, seed=21, n_threads=2)
    for _ in range(40):
        rec_a = a.collect()
        rec_b, off, lab = b.collect_csr()
        assert np.array_equal(rec_a, rec_b) and off[0] == 0 and len(off) == len(rec_b) + 1 and len(lab) == off[-1]
        if len(rec_a) == 0:
            break
        pol, val = _fake_net(rec_a)
        priors = np.zeros(len(lab), np.float64)
        for i in range(len(rec_b)):
            idx = lab[off[i]:off[i + 1]]
            raw = np.where(idx >= 0, pol[i][np.maximum(idx, 0)], 0.0).astype(np.float64)
            tot = 1e-8
            for x in raw:
                tot += x
            priors[off[i]:off[i + 1]] = raw / tot
        a.apply(pol, val)
        b.apply_priors(priors, val)
    assert a.stats() == b.stats() and a.stats()["moves_played"] > 0
    assert [a.fen(g) for g in range(4)] == [b.fen(g) for g in range(4)]
    for g in range(4):
        ea, eb = a.root_edges(g), b.root_edges(g)
        assert (ea is None) == (eb is None)
        if ea is not None:
            assert all(np.array_equal(x, y) for x, y in zip(ea, eb))
    a.close()
    b.close()
