"""On-device self-play driver and the trajectory exchange of SURVEY.md 8e.

``SelfPlay`` replaces ``Game.batch_self_play`` + ``AlphaZeroPlayer.get_batch_action`` (src/game.py:65-164,
src/player.py:333-375): every ply is search (device-resident playout loop) -> counts/root stats -> one kernel that
builds the policy target, samples the move, records the position and steps the env -> re-root -> one kernel that
turns finished games into packed training records and restarts their slots.  Nothing but a 4-byte counter crosses
PCIe per ply.  ``unpack_records`` returns the reference's tuples; ``all_gather_records`` is the single NCCL
all-gather of finished trajectories.  PyTorch = device memory, streams, torch.distributed only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, device_search as ds, mcts_cpp

_G = {"Connect4": (0, 6, 7, 7, 42), "Othello": (1, 8, 8, 65, 128)}


class AzSelfplay(C.Structure):
    _fields_ = [("game", C.c_int32), ("n", C.c_int32), ("max_plies", C.c_int32), ("td_steps", C.c_int32),
                ("temp_decay_moves", C.c_int32), ("temp_init", C.c_float), ("temp_endgame", C.c_float), ("seed", C.c_uint64),
                ("uid_stride", C.c_uint64), ("states", C.c_void_p), ("steps", C.c_void_p), ("uids", C.c_void_p),
                ("st_state", C.c_void_p), ("st_prob", C.c_void_p), ("st_wdl", C.c_void_p), ("st_mask", C.c_void_p),
                ("st_player", C.c_void_p), ("actions", C.c_void_p), ("finished", C.c_void_p), ("out", C.c_void_p),
                ("out_count", C.c_void_p), ("out_capacity", C.c_int32), ("record_bytes", C.c_int32)]


class AzLayout(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("record_bytes", "T1", "off_header", "off_state", "off_prob", "off_root_wdl", "off_future",
                                         "off_winner", "off_steps", "off_aux", "off_mask")]


def record_layout(game: str) -> AzLayout:
    L = AzLayout()
    if _lib.lib().az_selfplay_layout_for(_G[game][0], C.byref(L)) != 0:
        raise RuntimeError("az_selfplay_layout_for failed")
    return L


def shard_range(total: int, rank: int, world: int):
    """Contiguous game slice of rank r: [r*G/W, (r+1)*G/W)."""
    return (total * rank) // world, (total * (rank + 1)) // world


class SelfPlay:
    def __init__(self, game, n_slots, n_playout, vl_batch, evaluator, search_cfg=None, temperature=1.0, temp_decay_moves=20,
                 temp_endgame=0.0, td_steps=10, seed=0, uid_base=0, uid_stride=None, device=None, out_capacity=None, cache_size=0):
        if not torch.cuda.is_available():
            raise RuntimeError("SelfPlay needs a CUDA device (no CPU fallback)")
        self.game = game
        self.gid, self.R, self.Cc, self.A, self.T = _G[game]
        self.S = self.R * self.Cc
        self.n, self.n_playout, self.K = int(n_slots), int(n_playout), int(vl_batch)
        dev_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", dev_index)
        self.eval_cache = None
        if not isinstance(evaluator, ds.SyntheticEvaluator) and hasattr(evaluator, "predict_device"):
            if cache_size > 0:                              # device evaluation cache (src/Cache.py semantics)
                self.eval_cache = ds.EvalCache(game, cache_size, dev_index)
                evaluator = ds.CachedNetEvaluator(evaluator, self.eval_cache)
            else:
                evaluator = ds.NetEvaluator(evaluator)      # a network with the device contract
        self.evaluator = evaluator
        self.engine = getattr(mcts_cpp, f"BatchedMCTS_{game}")(self.n, device=dev_index)
        for k, v in (search_cfg or {}).items():
            setattr(self.engine.config, k, v)
        self.engine.set_seed(seed)
        self.engine.set_env_base(uid_base)
        self.layout = record_layout(game)
        d = dict(device=self.device)
        n, T, A, S = self.n, self.T, self.A, self.S
        self.out_capacity = int(out_capacity or 2 * n)
        self.states = torch.zeros((n, 32), dtype=torch.uint8, **d)
        self.steps = torch.zeros(n, dtype=torch.int32, **d)
        self.uids = (torch.arange(n, dtype=torch.int64, **d) + int(uid_base))
        self.st_state = torch.zeros((n, T, 3 * S), dtype=torch.int8, **d)
        self.st_prob = torch.zeros((n, T, A), dtype=torch.float32, **d)
        self.st_wdl = torch.zeros((n, T, 3), dtype=torch.float32, **d)
        self.st_mask = torch.zeros((n, T, A), dtype=torch.uint8, **d)
        self.st_player = torch.zeros((n, T), dtype=torch.int8, **d)
        self.actions = torch.zeros(n, dtype=torch.int32, **d)
        self.finished = torch.zeros(n, dtype=torch.uint8, **d)
        self.out = torch.zeros((self.out_capacity, self.layout.record_bytes), dtype=torch.uint8, **d)
        self.out_count = torch.zeros(1, dtype=torch.int32, **d)
        self.counts = torch.zeros((n, A), dtype=torch.int32, **d)
        self.stats = torch.zeros((n, 6 + 8 * A), dtype=torch.float32, **d)
        self.buf = ds.LeafBuffers(n, n * max(self.K, 1), A, (self.R, self.Cc), self.device,
                                  unpacked=not isinstance(self.evaluator, ds.SyntheticEvaluator),
                                  planes=not isinstance(self.evaluator, ds.SyntheticEvaluator))
        self.buf.roots = self.states                     # the env states ARE the search roots
        self.sp = AzSelfplay(self.gid, n, T, int(td_steps), int(temp_decay_moves), float(temperature), float(temp_endgame), int(seed),
                             int(uid_stride or n), self.states.data_ptr(), self.steps.data_ptr(), self.uids.data_ptr(),
                             self.st_state.data_ptr(), self.st_prob.data_ptr(), self.st_wdl.data_ptr(), self.st_mask.data_ptr(),
                             self.st_player.data_ptr(), self.actions.data_ptr(), self.finished.data_ptr(), self.out.data_ptr(),
                             self.out_count.data_ptr(), self.out_capacity, self.layout.record_bytes)
        self._L = _lib.lib()
        rc = self._L.az_envs_reset_dev(self.gid, n, self.states.data_ptr(), self._stream())
        if rc != 0:
            raise RuntimeError("az_envs_reset_dev failed")
        self.plies = 0
        self.launches = 0

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream or None

    def ply(self):
        """One lockstep ply of every slot (all launches asynchronous on the current stream)."""
        s = self._stream()
        self.launches += ds.playout_device(self.engine, self.buf, self.n_playout, self.K, self.evaluator, s or 0)
        self.engine.get_counts_dev(self.counts.data_ptr(), s or 0)
        self.engine.get_root_stats_dev(self.stats.data_ptr(), s or 0)
        if self._L.az_selfplay_ply_dev(C.byref(self.sp), self.counts.data_ptr(), self.stats.data_ptr(), s) != 0:
            raise RuntimeError("az_selfplay_ply_dev failed")
        self.engine.prune_roots_dev(self.actions.data_ptr(), s or 0)
        if self._L.az_selfplay_flush_dev(C.byref(self.sp), s) != 0:
            raise RuntimeError("az_selfplay_flush_dev failed")
        self.launches += 5
        self.plies += 1

    def run(self, target_games: int, max_plies: int | None = None):
        """Play until `target_games` games have finished.  Returns (records uint8[m, record_bytes] on the device, m)."""
        target_games = min(int(target_games), self.out_capacity)
        done = 0
        while done < target_games and (max_plies is None or self.plies < max_plies):
            self.ply()
            done = int(self.out_count.item())                      # the only host<->device traffic per ply: 4 bytes
        m = min(done, self.out_capacity)
        return self.out[:m], m

    @property
    def simulations(self):
        return self.plies * self.n * self.n_playout


def unpack_records(packed, game: str, td_steps: int = 1):
    """packed uint8[m, record_bytes] (numpy or tensor) -> list of dicts; ``tuples`` is what the reference's
    ``batch_self_play`` returns per game: (winner, tuple_of_training_tuples) (src/game.py:128-157)."""
    if isinstance(packed, torch.Tensor):
        packed = packed.cpu().numpy()
    L = record_layout(game)
    gid, R, Cc, A, T = _G[game]
    S = R * Cc
    games = []
    for rec in packed:
        length, winner = (int(x) for x in rec[L.off_header:L.off_header + 8].view(np.int32))
        uid = int(rec[L.off_header + 8:L.off_header + 16].view(np.uint64)[0])
        state = rec[L.off_state:L.off_state + L.T1 * 3 * S].view(np.int8).reshape(L.T1, 3, R, Cc)[:length]
        prob = rec[L.off_prob:L.off_prob + L.T1 * A * 4].view(np.float32).reshape(L.T1, A)[:length]
        wdl = rec[L.off_root_wdl:L.off_root_wdl + L.T1 * 12].view(np.float32).reshape(L.T1, 3)[:length]
        fut = rec[L.off_future:L.off_future + L.T1 * 12].view(np.float32).reshape(L.T1, 3)[:length]
        wz = rec[L.off_winner:L.off_winner + L.T1].view(np.int8)[:length].astype(np.int32)
        ste = rec[L.off_steps:L.off_steps + L.T1 * 2].view(np.int16)[:length].astype(np.int32)
        aux = rec[L.off_aux:L.off_aux + L.T1 * 2].view(np.int16)[:length].astype(np.int32)
        mask = rec[L.off_mask:L.off_mask + L.T1 * A].reshape(L.T1, A)[:length].astype(bool)
        rows = []
        for t in range(length):
            row = [state[t], prob[t], int(wz[t]), int(ste[t]), int(aux[t]), wdl[t], mask[t]]
            if td_steps > 0:
                row.append(fut[t])
            rows.append(tuple(row))
        games.append(dict(uid=uid, winner=winner, length=length, state=state, prob=prob, root_wdl=wdl, future_root_wdl=fut,
                          winner_z=wz, steps_to_end=ste, aux=aux, valid_mask=mask, tuples=(winner, tuple(rows))))
    return games


def all_gather_records(local: torch.Tensor, count: int, capacity: int):
    """The one collective of the path: every rank contributes `capacity` fixed-size record slots (the
    first `count` are valid) and receives everyone's.  Returns (records uint8[sum(counts), record_bytes], counts)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rb = local.shape[1]
    if world == 1:
        return local[:count], [count]
    send = torch.zeros((capacity, rb), dtype=torch.uint8, device=local.device)
    send[:count] = local[:count]
    recv = torch.empty((world * capacity, rb), dtype=torch.uint8, device=local.device)
    cnt = torch.tensor([count], dtype=torch.int32, device=local.device)
    cnts = torch.empty(world, dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(recv, send)
    dist.all_gather_into_tensor(cnts, cnt)
    counts = [int(c) for c in cnts.cpu()]
    parts = [recv[r * capacity:r * capacity + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0), counts


# ---------------------------------------------------------------------------------------------------------------
# Trajectory wire / disk formats: the reference's own layouts, produced from packed records.
# ---------------------------------------------------------------------------------------------------------------
def to_replay_tensors(packed: torch.Tensor, game: str):
    """Packed records (uint8[m, record_bytes], any device) -> the 8 tensors of src/ReplayBuffer.py:12-19, one row per
    recorded position (terminal tuples included), built with vectorised slicing on the tensor's own device."""
    L = record_layout(game)
    gid, R, Cc, A, T = _G[game]
    S, m, T1 = R * Cc, packed.shape[0], L.T1
    pk = packed.contiguous()
    length = pk[:, L.off_header:L.off_header + 4].view(torch.int32).reshape(m)
    keep = (torch.arange(T1, device=pk.device)[None, :] < length[:, None]).reshape(-1)

    def field(off, nbytes, dtype, shape):
        return pk[:, off:off + nbytes].contiguous().view(dtype).reshape(m * T1, *shape)[keep]

    return {
        "state": field(L.off_state, T1 * 3 * S, torch.int8, (3, R, Cc)),
        "prob": field(L.off_prob, T1 * A * 4, torch.float32, (A,)),
        "winner": field(L.off_winner, T1, torch.int8, (1,)),
        "steps_to_end": field(L.off_steps, T1 * 2, torch.int16, (1,)),
        "aux_target": field(L.off_aux, T1 * 2, torch.int16, (1,)),
        "root_wdl": field(L.off_root_wdl, T1 * 12, torch.float32, (3,)),
        "valid_mask": field(L.off_mask, T1 * A, torch.uint8, (A,)).bool(),
        "future_root_wdl": field(L.off_future, T1 * 12, torch.float32, (3,)),
    }


def save_replay_pt(path: str, tensors: dict):
    """Write the `.pt` layout that src/ReplayBuffer.py:25-62 saves/loads (state dict + `_ptr` + `current_capacity`)."""
    n = tensors["state"].shape[0]
    sd = {k: v.cpu() for k, v in tensors.items()}
    sd["_ptr"], sd["current_capacity"] = n, n
    torch.save(sd, path)


def to_upload_payload(games) -> bytes:
    """The pickle an actor POSTs to the learner's /upload endpoint (client.py:367-373): {'__az__': True, 'data': [play_data, ...]}
    where play_data is the tuple of training tuples of one game (`unpack_records(...)[i]['tuples'][1]`)."""
    import pickle
    return pickle.dumps({"__az__": True, "data": [g["tuples"][1] for g in games]}, protocol=pickle.HIGHEST_PROTOCOL)
