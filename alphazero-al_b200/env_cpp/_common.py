"""Shared implementation of the Connect4 / Othello ``Env`` classes: host-side mirror of the reference's pybind Env
objects (src/cpp/env_common.h:133-249).  The state is one 32-byte ``az_root`` bitboard record; every method is one
call into the C ABI (include/azb200_env.h)."""
from __future__ import annotations

import ctypes as C
import random

import numpy as np

from .. import _lib


class BitboardEnv:
    _GAME = -1
    _R = _C = _A = 0
    NUM_SYMMETRIES = 0
    __slots__ = ("_s",)

    def __init__(self, board=None):
        self._s = _lib.AzRoot()
        _lib.lib().az_env_reset(self._GAME, C.byref(self._s))
        if board is not None:
            self._set_board(board)

    # -- board I/O (env_common.h:34-81) -----------------------------------------------------------------------
    def _set_board(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.ndim != 2 or a.shape != (self._R, self._C):
            raise RuntimeError(f"board shape must be ({self._R}, {self._C})")
        b = a.astype(np.int8)
        L = _lib.lib()
        L.az_env_import(self._GAME, C.byref(self._s), b.ctypes.data_as(C.c_void_p))
        self._s.turn = 1 if L.az_env_n_pieces(self._GAME, C.byref(self._s)) % 2 == 0 else -1   # env_common.h:69

    @property
    def board(self):
        out = np.empty((self._R, self._C), np.int8)
        _lib.lib().az_env_export(self._GAME, C.byref(self._s), out.ctypes.data_as(C.c_void_p))
        return out.astype(np.float32)            # the reference returns a float32 copy (env_common.h:34-50)

    @board.setter
    def board(self, arr):
        self._set_board(arr)

    @property
    def turn(self):
        return int(self._s.turn)

    @turn.setter
    def turn(self, t):
        self._s.turn = int(t)

    # -- game logic ---------------------------------------------------------------------------------------------
    def reset(self):
        _lib.lib().az_env_reset(self._GAME, C.byref(self._s))

    def copy(self):
        e = type(self).__new__(type(self))
        e._s = _lib.AzRoot()
        C.memmove(C.byref(e._s), C.byref(self._s), C.sizeof(_lib.AzRoot))
        return e

    def step(self, action):
        _lib.lib().az_env_step(self._GAME, C.byref(self._s), int(action))

    def winPlayer(self):
        return _lib.lib().az_env_winner(self._GAME, C.byref(self._s))

    check_winner = winPlayer

    def check_full(self):
        return bool(_lib.lib().az_env_full(self._GAME, C.byref(self._s)))

    def done(self):
        return bool(_lib.lib().az_env_done(self._GAME, C.byref(self._s)))

    def valid_move(self):
        m = (C.c_int32 * self._A)()
        n = _lib.lib().az_env_valid_moves(self._GAME, C.byref(self._s), m)
        return [int(m[i]) for i in range(n)]

    def valid_mask(self):
        mask = [False] * self._A
        for a in self.valid_move():
            mask[a] = True
        return mask

    def current_state(self):
        """(1, 3, R, C) float32: own stones, opponent stones, side to move (env_common.h:93-119)."""
        b = self.board
        t = float(self._s.turn)
        st = np.zeros((1, 3, self._R, self._C), np.float32)
        st[0, 0] = b == t
        st[0, 1] = b == -t
        st[0, 2] = t
        return st

    def apply_symmetry(self, sym_id, inplace=False):
        e = self if inplace else self.copy()
        _lib.lib().az_env_apply_symmetry(self._GAME, C.byref(e._s), int(sym_id))
        return e

    def random_symmetry(self):
        sym = random.randrange(self.NUM_SYMMETRIES)        # all NUM_SYMMETRIES, own RNG (env_common.h:171-179)
        return self.apply_symmetry(sym), sym

    @classmethod
    def inverse_symmetry_action(cls, sym_id, action):
        return _lib.lib().az_env_inverse_symmetry_action(cls._GAME, int(sym_id), int(action))

    # -- pickle: (board float32, turn) (env_common.h:236-248) ------------------------------------------------
    def __getstate__(self):
        return (self.board, self.turn)

    def __setstate__(self, st):
        if len(st) != 2:
            raise RuntimeError("Invalid pickle state")
        self._s = _lib.AzRoot()
        _lib.lib().az_env_reset(self._GAME, C.byref(self._s))
        self._set_board(st[0])
        self._s.turn = int(st[1])

    def _render(self, header, footer, empty, row_prefix=False):
        b = self.board
        lines = list(header)
        for r in range(self._R):
            cells = " ".join(empty if v == 0 else ("X" if v == 1 else "O") for v in b[r])
            lines.append((f"{r} " if row_prefix else "") + cells)
        lines += list(footer)
        return "\n".join(lines)
