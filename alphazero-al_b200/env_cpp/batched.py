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
