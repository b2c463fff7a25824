"""Drop-in for the reference's `core` package (core/__init__.py): the same names, backed by the
pybind11 module CorePyExt of this repo (host C++17 mirror of Board / Policy / MCTS whose simulate
slots run on the B200).  `from gomokuai_b200.core import Board, MCTS, RandomPolicy, ...`"""
import os

from . import build as _build

if not os.path.exists(_build.ext_path()):
    raise ImportError(f"{_build.ext_path()} is missing: run `python -m gomokuai_b200.build`")

from .CorePyExt import GameConfig, Player, Position, Board          # noqa: E402,F401
from .CorePyExt import Node, Policy, MCTS                           # noqa: E402,F401
from .CorePyExt import RandomPolicy, PoolRAVEPolicy, TraditionalPolicy   # noqa: E402,F401
from .CorePyExt import RootParallelSearch, init, seed               # noqa: E402,F401

__doc__ = f"C++ extension 'core' (B200 hot path) at '{_build.ext_path()}'"
