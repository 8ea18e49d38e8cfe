"""Import alias: the package directory is ``chess-alpha-zero_b200/`` (the name the build contract
fixes); a hyphen is not importable, so ``import chess_alpha_zero_b200`` resolves here and this module
replaces itself in ``sys.modules`` with the real package."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chess-alpha-zero_b200")
_spec = importlib.util.spec_from_file_location(
    "chess_alpha_zero_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["chess_alpha_zero_b200"] = _mod
_spec.loader.exec_module(_mod)
