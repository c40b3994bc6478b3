"""``env_cpp.connect4.Env`` - same surface as the reference (src/cpp/env_connect4.h:20-66)."""
from ._common import BitboardEnv


class Env(BitboardEnv):
    _GAME, _R, _C, _A = 0, 6, 7, 7
    NUM_SYMMETRIES = 2
    __slots__ = ()

    def show(self):
        print(self._render(["===================="], ["0 1 2 3 4 5 6", "===================="], "_"))
