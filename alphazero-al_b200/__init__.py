"""alphazero-al_b200: B200-native drop-in for the self-play hot path of Sunshine-718/AlphaZero-AL.

The directory name carries a hyphen (it mirrors the upstream repo name); import it with
``importlib.import_module("alphazero-al_b200")`` or through the ``alphazero_al_b200`` alias module at the
repository root.

    mcts_cpp    - same Python surface as the reference's pybind module src.mcts_cpp
    evaluators  - deterministic synthetic leaf evaluators (numpy twins of csrc/az_eval.cu)
"""
from . import _lib  # noqa: F401


def build(force: bool = False):
    return _lib.build(force)
