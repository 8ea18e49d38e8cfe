"""The oracle restatements vs. fixtures produced by executing the reference itself
(oracle/pin_against_reference.py).  CPU only."""
import hashlib
import json
import os

import numpy as np

from oracle import encoder, labels, puct


def test_labels_match_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "labels.json")))
    assert labels.N_LABELS == 1968 == len(g["labels"])
    assert labels.LABELS == g["labels"]
    assert labels.flipped_labels() == g["flipped_labels"]
    assert labels.UNFLIPPED_INDEX.tolist() == g["unflipped_index"]
    # hashes recorded in SURVEY.md 8(a10)
    assert hashlib.sha256(",".join(labels.LABELS).encode()).hexdigest().startswith("409054c3")
    assert hashlib.sha256(labels.UNFLIPPED_INDEX.astype("<i4").tobytes()).hexdigest().startswith("deb182ab")
    for lab, idx in (("e2e4", 930), ("e7e5", 1097), ("g1f3", 1402), ("e1g1", 901), ("a7a8q", 1793)):
        assert labels.LABEL_INDEX[lab] == idx


def test_unflipped_index_is_an_involution_without_fixed_points():
    u = labels.UNFLIPPED_INDEX
    assert np.array_equal(u[u], np.arange(1968))
    assert not np.any(u == np.arange(1968))


def test_flip_policy_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "flip_policy.npz"))
    for p, f in zip(g["policy"], g["flipped"]):
        assert np.array_equal(labels.flip_policy(p), f)


def test_encoder_matches_reference_bit_exact(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "encoder_fens.json")))
    want = np.load(os.path.join(golden_dir, "encoder.npz"))["planes_u8"].astype(np.float32)
    assert len(meta["fens"]) == 261
    for fen, w, h in zip(meta["fens"], want, meta["sha256_16"]):
        got = encoder.canon_input_planes(fen)
        assert got.dtype == np.float32 and got.shape == (18, 8, 8)
        assert got.tobytes() == w.tobytes(), fen
        assert hashlib.sha256(got.tobytes()).hexdigest()[:16] == h
        # the compact board record (input of the CUDA encoder) carries the same information
        assert encoder.record_to_planes(encoder.fen_to_record(fen)).tobytes() == w.tobytes(), fen


def test_encoder_known_answers_from_survey():
    # sha256 of the C-order float32 bytes, first 16 hex digits
    want = ["ac3c7e683e34b11f", "6c2256c70c78833e", "322f8d24f0d2be4a", "99c21d6932b5d66d", "f005cc15c3db0418"]
    for fen, h in zip(encoder.KNOWN_ANSWER_FENS, want):
        assert hashlib.sha256(encoder.canon_input_planes(fen).tobytes()).hexdigest()[:16] == h
    p = encoder.canon_input_planes(encoder.KNOWN_ANSWER_FENS[1])
    assert p[17, 5, 4] == 1 and p[17].sum() == 1          # ep plane is NOT mirrored for black
    p = encoder.canon_input_planes(encoder.KNOWN_ANSWER_FENS[3])
    assert [p[12 + i, 0, 0] for i in range(4)] == [0, 1, 1, 0] and p[16, 3, 3] == 37.0


def test_puct_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "puct.npz"))
    off = g["offsets"]
    c_puct, eps = float(g["c_puct"]), float(g["noise_eps"])
    assert (c_puct, eps) == (1.5, 0.25)
    ties = 0
    for i in range(len(off) - 1):
        s = slice(off[i], off[i + 1])
        pushed = puct.push_priors_from_legal(g["policy_legal"][s])
        assert np.array_equal(pushed, g["pushed"][s])       # bit-exact float64
        noise = g["noise"][s] if g["is_root"][i] else None
        best = puct.select(g["n"][s], g["q"][s], g["p"][s], int(g["sum_n"][i]), c_puct, noise, eps)
        assert best == int(g["best"][i]), i
        ties += len(set(g["p"][s].tolist())) == 1 and off[i + 1] - off[i] >= 3
    assert ties >= 10      # the exact-tie cases (first maximum wins) are really in the fixture
