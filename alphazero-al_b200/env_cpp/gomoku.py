"""``env_cpp.gomoku.Env`` - Env-only API parity with the reference's byte-board Gomoku (src/cpp/Gomoku.h:11-296,
src/cpp/env_gomoku.h:60-171).  The reference registers no MCTS engine for Gomoku (mcts_bindings.cpp:393-394).  The
state is one 288-byte ``az_gomoku`` record of row bit masks; every game-logic method is one call into the C ABI
(include/azb200_gomoku.h) - the same code the lockstep device kernels run (``BatchedGomoku``)."""
from __future__ import annotations

import ctypes as C
import random

import numpy as np

from .. import _lib

_MSG = {1: "game is already finished", 2: "action out of range", 3: "cell is already occupied",
        4: "board_size must be positive", 5: "n_in_row must be >= 2", 6: "n_in_row must be <= board size",
        7: "board values must be -1, 0, or 1", 8: "invalid symmetry id",
        9: "board_size > 32 is not supported by this implementation (32-bit row masks)"}


def _ck(rc):
    if rc != 0:
        raise RuntimeError(_MSG.get(abs(rc), f"gomoku error {rc}"))


class Env:
    NUM_SYMMETRIES = 8
    __slots__ = ("_s",)

    def __init__(self, board_size=15, n_in_row=5, board=None):
        self._s = _lib.AzGomoku()
        if board is None and not isinstance(board_size, (int, np.integer)):
            board, board_size = board_size, None           # Env(board, n_in_row=5) overload (env_gomoku.h:70-73)
        if board is not None:
            a = np.asarray(board, dtype=np.float32)
            if a.ndim != 2 or a.shape[0] != a.shape[1]:
                raise RuntimeError("board must be square")
            self.set_params(int(a.shape[0]), int(n_in_row))
            self.board = a
        else:
            self.set_params(int(board_size), int(n_in_row))

    def _p(self):
        return C.byref(self._s)

    # -- configuration (Gomoku.h:21-28, 214-222) ---------------------------------------------------------------
    def set_params(self, board_size, n_in_row):
        _ck(_lib.lib().az_gomoku_set_params(self._p(), int(board_size), int(n_in_row)))

    def reset(self):
        _lib.lib().az_gomoku_reset(self._p())

    board_size = property(lambda self: int(self._s.size))
    rows = property(lambda self: int(self._s.size))
    cols = property(lambda self: int(self._s.size))
    n_in_row = property(lambda self: int(self._s.n_in_row))
    action_size = property(lambda self: int(self._s.size) ** 2)
    num_symmetries = property(lambda self: 8)

    @property
    def turn(self):
        return int(self._s.turn)

    @turn.setter
    def turn(self, t):
        if t != 1 and t != -1:
            raise RuntimeError("turn must be 1 or -1")
        self._s.turn = int(t)

    @property
    def board(self):
        n = self.board_size
        out = np.empty((n, n), np.int8)
        _lib.lib().az_gomoku_export(self._p(), out.ctypes.data_as(C.c_void_p))
        return out.astype(np.float32)            # the reference returns a float32 copy (env_gomoku.h:30-44)

    @board.setter
    def board(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        n = self.board_size
        if a.ndim != 2 or a.shape != (n, n):
            raise RuntimeError("board shape does not match environment dimensions")
        b = np.ascontiguousarray(a.astype(np.int8))
        _ck(_lib.lib().az_gomoku_import(self._p(), b.ctypes.data_as(C.c_void_p)))   # import_board + sync_from_board

    def step(self, action):                                    # Gomoku.h:63-92 (validated, unlike Connect4/Othello)
        _ck(_lib.lib().az_gomoku_step(self._p(), int(action)))

    def step_xy(self, row, col):
        self.step(self.coord_to_action(row, col))

    def coord_to_action(self, row, col):
        n = self.board_size
        if not (0 <= row < n and 0 <= col < n):
            raise RuntimeError("row/col out of range")
        return row * n + col

    def action_to_coord(self, action):
        if action < 0 or action >= self.action_size:
            raise RuntimeError("action out of range")
        return (action // self.board_size, action % self.board_size)

    def winPlayer(self):
        return int(self._s.winner)

    check_winner = winPlayer

    def check_full(self):
        return int(self._s.n_pieces) == self.action_size

    def done(self):
        return bool(self._s.done)

    def valid_move(self):
        m = np.empty(self.action_size, np.int32)
        k = _lib.lib().az_gomoku_valid_moves(self._p(), m.ctypes.data_as(C.c_void_p))
        return m[:k].tolist()

    def valid_mask(self):
        return [bool(v) for v in (self.board.reshape(-1) == 0)]

    def current_state(self):
        b, t, n = self.board, float(self.turn), self.board_size
        st = np.zeros((1, 3, n, n), np.float32)
        st[0, 0], st[0, 1], st[0, 2] = b == t, b == -t, t
        return st

    def copy(self):
        e = Env.__new__(Env)
        e._s = _lib.AzGomoku()
        C.memmove(C.byref(e._s), self._p(), C.sizeof(_lib.AzGomoku))
        return e

    def apply_symmetry(self, sym_id, inplace=False):
        e = self if inplace else self.copy()
        _ck(_lib.lib().az_gomoku_apply_symmetry(e._p(), int(sym_id)))
        return e

    def random_symmetry(self):
        sym = random.randrange(8)
        return self.apply_symmetry(sym), sym

    def inverse_symmetry_action(self, sym_id, action):         # Gomoku.h:115-128 (applies transform_coord(sym_id))
        r = _lib.lib().az_gomoku_inverse_symmetry_action(self.board_size, int(sym_id), int(action))
        if r < 0:
            _ck(r)
        return int(r)

    def show(self):
        n = self.board_size
        b = self.board.reshape(-1)
        lines = ["==============================", "    " + "".join(f"{c % 10} " for c in range(n))]
        for r in range(n):
            row = "".join(("." if v == 0 else ("X" if v == 1 else "O")) + " " for v in b[r * n:(r + 1) * n])
            lines.append((" " if r < 10 else "") + f"{r}  " + row)
        lines.append("==============================")
        print("\n".join(lines))

    def __getstate__(self):                                     # pickle = (board, turn, n_in_row) (env_gomoku.h:151-168)
        return (self.board, self.turn, self.n_in_row)

    def __setstate__(self, st):
        if len(st) != 3:
            raise RuntimeError("Invalid pickle state")
        a = np.asarray(st[0], dtype=np.float32)
        self._s = _lib.AzGomoku()
        self.set_params(int(a.shape[0]), int(st[2]))
        self.board = a
        self.turn = int(st[1])
