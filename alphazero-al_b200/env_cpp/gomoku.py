"""``env_cpp.gomoku.Env`` - Env-only API parity with the reference's byte-board Gomoku (src/cpp/Gomoku.h:11-296,
src/cpp/env_gomoku.h:60-171).  The reference registers no MCTS engine for Gomoku (mcts_bindings.cpp:393-394), so this
is host-side API glue: runtime board size, validated step, incremental line check, D4 symmetries, pickle."""
from __future__ import annotations

import random

import numpy as np


class Env:
    NUM_SYMMETRIES = 8
    _DIRS = ((1, 0), (0, 1), (1, 1), (1, -1))

    def __init__(self, board_size=15, n_in_row=5, board=None):
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

    # -- configuration (Gomoku.h:21-28, 214-222) ---------------------------------------------------------------
    def set_params(self, board_size, n_in_row):
        if board_size <= 0:
            raise RuntimeError("board_size must be positive")
        if n_in_row <= 1:
            raise RuntimeError("n_in_row must be >= 2")
        if n_in_row > board_size:
            raise RuntimeError("n_in_row must be <= board size")
        self._n, self._k = int(board_size), int(n_in_row)
        self._b = np.zeros(self._n * self._n, np.int8)
        self.reset()

    def reset(self):
        self._b[:] = 0
        self._turn, self._pieces, self._last_action, self._last_player, self._winner, self._done = 1, 0, -1, 0, 0, False

    board_size = property(lambda self: self._n)
    rows = property(lambda self: self._n)
    cols = property(lambda self: self._n)
    n_in_row = property(lambda self: self._k)
    action_size = property(lambda self: self._n * self._n)
    num_symmetries = property(lambda self: 8)

    @property
    def turn(self):
        return self._turn

    @turn.setter
    def turn(self, t):
        if t != 1 and t != -1:
            raise RuntimeError("turn must be 1 or -1")
        self._turn = int(t)

    @property
    def board(self):
        return self._b.reshape(self._n, self._n).astype(np.float32)

    @board.setter
    def board(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.ndim != 2 or a.shape != (self._n, self._n):
            raise RuntimeError("board shape does not match environment dimensions")
        self._b = a.astype(np.int8).reshape(-1).copy()
        self._sync_from_board()

    def _sync_from_board(self):                                # Gomoku.h:160-204
        b = self._b
        if np.any((b != 0) & (b != 1) & (b != -1)):
            raise RuntimeError("board values must be -1, 0, or 1")
        p1, p2 = int(np.sum(b == 1)), int(np.sum(b == -1))
        self._pieces = p1 + p2
        nz = np.nonzero(b)[0]
        self._last_action = int(nz[-1]) if nz.size else -1
        self._last_player = int(b[nz[-1]]) if nz.size else 0
        if p1 == p2:
            self._turn = 1
        elif p1 == p2 + 1:
            self._turn = -1
        else:
            self._turn = 1 if self._pieces % 2 == 0 else -1
        self._winner = 0
        for i in nz:                                           # find_winner_full_scan (Gomoku.h:265-274)
            if self._has_line_from(int(i), int(b[i])):
                self._winner = int(b[i])
                break
        self._done = self._winner != 0 or self._pieces == self.action_size

    def _count(self, r, c, dr, dc, player):
        n, cnt = self._n, 0
        r, c = r + dr, c + dc
        while 0 <= r < n and 0 <= c < n and self._b[r * n + c] == player:
            cnt += 1
            r, c = r + dr, c + dc
        return cnt

    def _has_line_from(self, action, player):                  # Gomoku.h:247-263
        r, c = divmod(action, self._n)
        return any(1 + self._count(r, c, dr, dc, player) + self._count(r, c, -dr, -dc, player) >= self._k for dr, dc in self._DIRS)

    def step(self, action):                                    # Gomoku.h:63-92 (validated, unlike Connect4/Othello)
        action = int(action)
        if self._done:
            raise RuntimeError("game is already finished")
        if action < 0 or action >= self.action_size:
            raise RuntimeError("action out of range")
        if self._b[action] != 0:
            raise RuntimeError("cell is already occupied")
        self._b[action] = self._turn
        self._pieces += 1
        self._last_action, self._last_player = action, self._turn
        if self._has_line_from(action, self._last_player):
            self._winner, self._done = self._last_player, True
        elif self._pieces == self.action_size:
            self._winner, self._done = 0, True
        self._turn = -self._turn

    def step_xy(self, row, col):
        self.step(self.coord_to_action(row, col))

    def coord_to_action(self, row, col):
        if not (0 <= row < self._n and 0 <= col < self._n):
            raise RuntimeError("row/col out of range")
        return row * self._n + col

    def action_to_coord(self, action):
        if action < 0 or action >= self.action_size:
            raise RuntimeError("action out of range")
        return (action // self._n, action % self._n)

    def winPlayer(self):
        return self._winner

    check_winner = winPlayer

    def check_full(self):
        return self._pieces == self.action_size

    def done(self):
        return self._done

    def valid_move(self):
        return [int(i) for i in np.nonzero(self._b == 0)[0]]

    def valid_mask(self):
        return [bool(v) for v in (self._b == 0)]

    def current_state(self):
        b, t = self.board, float(self._turn)
        st = np.zeros((1, 3, self._n, self._n), np.float32)
        st[0, 0], st[0, 1], st[0, 2] = b == t, b == -t, t
        return st

    def copy(self):
        e = Env.__new__(Env)
        e.__dict__.update(self.__dict__)
        e._b = self._b.copy()
        return e

    def _xform(self, sym, r, c):                               # Gomoku.h:276-294
        n = self._n
        if sym < 0 or sym >= 8:
            raise RuntimeError("invalid symmetry id")
        return ((r, c), (c, n - 1 - r), (n - 1 - r, n - 1 - c), (n - 1 - c, r), (r, n - 1 - c), (n - 1 - r, c), (c, r),
                (n - 1 - c, n - 1 - r))[sym]

    def apply_symmetry(self, sym_id, inplace=False):
        e = self if inplace else self.copy()
        sym_id = int(sym_id)
        if sym_id < 0 or sym_id >= 8:
            raise RuntimeError("invalid symmetry id")
        if sym_id == 0:
            return e
        n = e._n
        old = e._b.reshape(n, n)
        new = np.zeros_like(old)
        rr, cc = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        nr, nc = e._xform(sym_id, rr, cc)
        new[nr, nc] = old[rr, cc]
        e._b = new.reshape(-1).copy()
        if e._last_action >= 0:
            r, c = divmod(e._last_action, n)
            r2, c2 = e._xform(sym_id, r, c)
            e._last_action = int(r2 * n + c2)
        return e

    def random_symmetry(self):
        sym = random.randrange(8)
        return self.apply_symmetry(sym), sym

    def inverse_symmetry_action(self, sym_id, action):         # Gomoku.h:115-128 (applies transform_coord(sym_id))
        if action < 0 or action >= self.action_size:
            raise RuntimeError("action out of range")
        r, c = divmod(int(action), self._n)
        r2, c2 = self._xform(int(sym_id), r, c)
        return int(r2 * self._n + c2)

    def show(self):
        n = self._n
        lines = ["==============================", "    " + "".join(f"{c % 10} " for c in range(n))]
        for r in range(n):
            row = "".join(("." if v == 0 else ("X" if v == 1 else "O")) + " " for v in self._b[r * n:(r + 1) * n])
            lines.append((" " if r < 10 else "") + f"{r}  " + row)
        lines.append("==============================")
        print("\n".join(lines))

    def __getstate__(self):                                     # pickle = (board, turn, n_in_row) (env_gomoku.h:151-168)
        return (self.board, self._turn, self._k)

    def __setstate__(self, st):
        if len(st) != 3:
            raise RuntimeError("Invalid pickle state")
        a = np.asarray(st[0], dtype=np.float32)
        self.set_params(int(a.shape[0]), int(st[2]))
        self.board = a
        self.turn = int(st[1])
