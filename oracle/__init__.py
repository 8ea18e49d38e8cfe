"""CPU oracle for the chess-alpha-zero evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU arm.  The product package (``chess-alpha-zero_b200/``) never imports
this package and fails loudly if its CUDA library is missing.

What each module restates (reference paths relative to /root/reference):

  labels.py   src/chess_zero/config.py:76-123,143-146,175-185   (1968 UCI labels, flip index)
  encoder.py  src/chess_zero/env/chess_env.py:231-348            (FEN -> 18x8x8 planes)
  puct.py     src/chess_zero/agent/player_chess.py:251-299       (prior push + PUCT select)
  network.py  src/chess_zero/agent/model_chess.py:57-111         (Keras graph, fp32 CPU)
  weights.py  Keras default initialisers named in data/model/model_best_config.json

Pinning status
  * labels / encoder / puct: PINNED.  ``pin_against_reference.py`` imports the
    reference's own modules in the build container (chess_zero.config as-is;
    chess_env / player_chess under a stub ``chess`` module), runs them on seeded
    inputs and writes the results to ``tests/golden/``; ``tests/test_oracle_*.py``
    check the restatements against those fixtures bit-for-bit.
  * network forward: PARITY UNPINNED.  The arithmetic lives in un-vendored
    third-party dependencies (keras==2.0.8, tensorflow-gpu==1.15.2,
    requirements.txt:2-3) that are not installed and cannot be installed here,
    and the reference has no tests or golden tensors for it.  The restatement
    follows the published Keras semantics (channels_first cross-correlation with
    symmetric "same" padding, BatchNormalization inference formula with
    epsilon from the JSON, Flatten on NCHW, Dense x@W+b) and is cross-checked by
    two independent formulations (graph interpreter on un-folded NCHW tensors vs
    folded NHWC im2col GEMM) that must agree to <= 1e-5.
"""
