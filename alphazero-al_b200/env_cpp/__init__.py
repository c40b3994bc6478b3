"""Drop-in replacement for the reference's pybind module ``src.env_cpp`` (src/cpp/env_bindings.cpp:17-25): importable
submodules ``connect4``, ``othello``, ``gomoku``, each exporting ``Env`` (src/environments/<Game>/__init__.py:1 does
``from src.env_cpp.<game> import Env``).  ``BatchedEnv`` / ``BatchedGomoku`` (B200-only) advance N games in lockstep on the device."""
from . import connect4, gomoku, othello  # noqa: F401
from .batched import BatchedEnv, BatchedGomoku  # noqa: F401
