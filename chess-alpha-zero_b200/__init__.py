"""B200-native (sm_100a) batched policy/value evaluation for chess-alpha-zero.

One hot path of the reference, rebuilt: input-plane encoder -> conv stem -> residual tower ->
policy/value heads, behind the reference's own ``ChessModel`` / ``ChessModelAPI`` / pipe protocol,
plus PUCT select.  All compute is in ``csrc/`` (hand-written CUDA, tcgen05 + TMA) behind the C ABI in
``include/caz_b200.h``; this package is the Python host mirror of the reference's interface.
There is no CPU fallback: importing works anywhere, computing needs the built library and a B200.
"""
from . import _lib, encoder, keras_graph, labels, mcts
from ._lib import CazError, library_path
from .api import ChessModelAPI
from .config import Config
from .evaluator import Evaluator
from .mcts import SelfPlaySearch
from .model import ChessModel
from .trainer import TrainConfig, Trainer

__all__ = ["ChessModel", "ChessModelAPI", "Evaluator", "Config", "CazError", "library_path", "encoder",
           "keras_graph", "labels", "mcts", "SelfPlaySearch", "Trainer", "TrainConfig"]
