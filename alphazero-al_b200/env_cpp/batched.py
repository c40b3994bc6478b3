"""N games advanced in lockstep on the device (include/azb200_env.h: az_envs_*_dev).  The state tensor is an
``az_root[n]`` array in HBM - exactly what ``BatchedMCTS_*.search_dev`` takes as roots - so self-play never leaves
the GPU.  PyTorch provides the device memory only."""
from __future__ import annotations

import torch

from .. import _lib

_GAMES = {"Connect4": (0, 6, 7, 7, 42), "Othello": (1, 8, 8, 65, 128)}


class BatchedEnv:
    def __init__(self, game: str, n: int, device=None):
        self.game = game
        self.gid, self.R, self.C, self.A, self.max_plies = _GAMES[game]
        self.n = int(n)
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedEnv needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.states = torch.zeros((self.n, 32), dtype=torch.uint8, device=self.device)     # az_root[n]
        self._L = _lib.lib()
        self.reset()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream or None

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc})")

    def reset(self):
        self._ck(self._L.az_envs_reset_dev(self.gid, self.n, self.states.data_ptr(), self._stream()), "az_envs_reset_dev")

    def step(self, actions: torch.Tensor, winners: torch.Tensor | None = None, dones: torch.Tensor | None = None):
        """actions int32[n] on the device; finished games and negative actions are skipped."""
        assert actions.dtype == torch.int32 and actions.is_cuda and actions.numel() == self.n
        p = lambda t: t.data_ptr() if t is not None else None
        self._ck(self._L.az_envs_step_dev(self.gid, self.n, self.states.data_ptr(), actions.data_ptr(), p(winners), p(dones),
                                          self._stream()), "az_envs_step_dev")

    def observe(self, boards=True, masks=True):
        """Returns dict of device tensors: boards int8[n,R,C], masks u8[n,A], turns i32[n], winners i32[n], dones u8[n]."""
        d = dict(device=self.device)
        out = dict(boards=torch.empty((self.n, self.R, self.C), dtype=torch.int8, **d) if boards else None,
                   masks=torch.empty((self.n, self.A), dtype=torch.uint8, **d) if masks else None,
                   turns=torch.empty(self.n, dtype=torch.int32, **d), winners=torch.empty(self.n, dtype=torch.int32, **d),
                   dones=torch.empty(self.n, dtype=torch.uint8, **d))
        p = lambda t: t.data_ptr() if t is not None else None
        self._ck(self._L.az_envs_observe_dev(self.gid, self.n, self.states.data_ptr(), p(out["boards"]), p(out["masks"]),
                                             p(out["turns"]), p(out["winners"]), p(out["dones"]), self._stream()), "az_envs_observe_dev")
        return out

    def random_rollouts(self, seed: int, first_game: int = 0, n_record: int = 0):
        """Config-2 workload: every game plays hash-chosen legal moves to the end on the device."""
        d = dict(device=self.device)
        digest = torch.empty(self.n, dtype=torch.int64, **d)
        plies = torch.empty(self.n, dtype=torch.int32, **d)
        rec = None
        ptrs = [None] * 6
        if n_record > 0:
            mp = self.max_plies
            rec = dict(boards=torch.zeros((n_record, mp, self.R, self.C), dtype=torch.int8, **d),
                       masks=torch.zeros((n_record, mp, self.A), dtype=torch.uint8, **d),
                       turns=torch.zeros((n_record, mp), dtype=torch.int32, **d), actions=torch.zeros((n_record, mp), dtype=torch.int32, **d),
                       winners=torch.zeros((n_record, mp), dtype=torch.int32, **d), dones=torch.zeros((n_record, mp), dtype=torch.uint8, **d))
            ptrs = [rec[k].data_ptr() for k in ("boards", "masks", "turns", "actions", "winners", "dones")]
        self._ck(self._L.az_envs_rollout_dev(self.gid, self.n, seed, first_game, digest.data_ptr(), plies.data_ptr(), n_record,
                                             self.max_plies, *ptrs, self._stream()), "az_envs_rollout_dev")
        return digest, plies, rec


class BatchedGomoku:
    """N Gomoku games advanced in lockstep on the device (include/azb200_gomoku.h: az_gomoku_*_dev).  The state tensor is
    an ``az_gomoku[n]`` array (288-byte records of row bit masks) in HBM."""
    RECORD_BYTES = 288

    def __init__(self, n: int, board_size: int = 15, n_in_row: int = 5, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedGomoku needs a CUDA device (no CPU fallback)")
        self.n, self.size, self.k = int(n), int(board_size), int(n_in_row)
        self.S = self.A = self.size * self.size
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.states = torch.zeros((self.n, self.RECORD_BYTES), dtype=torch.uint8, device=self.device)   # az_gomoku[n]
        self._L = _lib.lib()
        self.reset()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream or None

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc})")

    def reset(self):
        self._ck(self._L.az_gomoku_reset_dev(self.n, self.size, self.k, self.states.data_ptr(), self._stream()), "az_gomoku_reset_dev")

    def step(self, actions: torch.Tensor, status: torch.Tensor | None = None, winners: torch.Tensor | None = None,
             dones: torch.Tensor | None = None):
        """actions int32[n] on the device; finished games and negative actions are skipped, illegal actions are reported in
        ``status`` (u8[n]: 2 = out of range, 3 = occupied) and leave the game untouched."""
        assert actions.dtype == torch.int32 and actions.is_cuda and actions.numel() == self.n
        p = lambda t: t.data_ptr() if t is not None else None
        self._ck(self._L.az_gomoku_step_dev(self.n, self.states.data_ptr(), actions.data_ptr(), p(status), p(winners), p(dones),
                                            self._stream()), "az_gomoku_step_dev")

    def observe(self, boards=True, masks=True):
        """dict of device tensors: boards int8[n,S,S], masks u8[n,S*S], turns i32[n], winners i32[n], dones u8[n]."""
        d = dict(device=self.device)
        out = dict(boards=torch.empty((self.n, self.size, self.size), dtype=torch.int8, **d) if boards else None,
                   masks=torch.empty((self.n, self.A), dtype=torch.uint8, **d) if masks else None,
                   turns=torch.empty(self.n, dtype=torch.int32, **d), winners=torch.empty(self.n, dtype=torch.int32, **d),
                   dones=torch.empty(self.n, dtype=torch.uint8, **d))
        p = lambda t: t.data_ptr() if t is not None else None
        self._ck(self._L.az_gomoku_observe_dev(self.n, self.size, self.states.data_ptr(), p(out["boards"]), p(out["masks"]),
                                               p(out["turns"]), p(out["winners"]), p(out["dones"]), self._stream()),
                 "az_gomoku_observe_dev")
        return out

    def apply_symmetry(self, sym_ids: torch.Tensor):
        assert sym_ids.dtype == torch.int32 and sym_ids.is_cuda and sym_ids.numel() == self.n
        self._ck(self._L.az_gomoku_symmetry_dev(self.n, self.states.data_ptr(), sym_ids.data_ptr(), self._stream()),
                 "az_gomoku_symmetry_dev")

    def random_rollouts(self, seed: int, first_game: int = 0, n_record: int = 0, keep_final: bool = True):
        """Every game plays hash-chosen empty cells to the end on the device; the final records replace ``states``."""
        d = dict(device=self.device)
        digest = torch.empty(self.n, dtype=torch.int64, **d)
        plies = torch.empty(self.n, dtype=torch.int32, **d)
        rec, ptrs = None, [None] * 5
        if n_record > 0:
            S = self.S
            rec = dict(boards=torch.zeros((n_record, S, self.size, self.size), dtype=torch.int8, **d),
                       turns=torch.zeros((n_record, S), dtype=torch.int32, **d), actions=torch.zeros((n_record, S), dtype=torch.int32, **d),
                       winners=torch.zeros((n_record, S), dtype=torch.int32, **d), dones=torch.zeros((n_record, S), dtype=torch.uint8, **d))
            ptrs = [rec[k].data_ptr() for k in ("boards", "turns", "actions", "winners", "dones")]
        self._ck(self._L.az_gomoku_rollout_dev(self.n, self.size, self.k, seed, first_game, digest.data_ptr(), plies.data_ptr(),
                                               n_record, *ptrs, self.states.data_ptr() if keep_final else None, self._stream()),
                 "az_gomoku_rollout_dev")
        return digest, plies, rec
