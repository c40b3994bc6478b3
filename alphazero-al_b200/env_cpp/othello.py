"""``env_cpp.othello.Env`` - same surface as the reference (src/cpp/env_othello.h:20-74)."""
from ._common import BitboardEnv


class Env(BitboardEnv):
    _GAME, _R, _C, _A = 1, 8, 8, 65
    NUM_SYMMETRIES = 8
    __slots__ = ()

    def show(self):
        print(self._render(["========================", "  0 1 2 3 4 5 6 7"], ["========================"], ".", row_prefix=True))
