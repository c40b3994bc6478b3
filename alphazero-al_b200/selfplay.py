"""On-device self-play driver, the trajectory exchange of SURVEY.md 8e and the reference's wire / disk
formats (8f row 4).

``SelfPlay`` replaces ``Game.batch_self_play`` + ``AlphaZeroPlayer.get_batch_action`` (src/game.py:65-164,
src/player.py:333-375): every ply is search (device-resident playout loop) -> counts / root stats -> one kernel that builds the
policy target, samples the move, records the position and steps the env -> re-root -> one kernel that moves finished
trajectories into an output ring and restarts their slots.  Finished games are COMPACT records (include/azb200_selfplay.h:
32-byte game header + one 64-byte position record per position for Connect4); ``Records.to_replay_tensors`` expands them on the
device into the reference's training tuples / replay-buffer tensors, ``TrajectoryExchange`` is the single NCCL all-gather of the
path, issued on a side stream so that it overlaps the next batch's search.  PyTorch = device memory, streams, torch.distributed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, device_search as ds, mcts_cpp

_G = {"Connect4": (0, 6, 7, 7), "Othello": (1, 8, 8, 65)}
GAME_BYTES = 32


class AzSelfplay(C.Structure):
    """az_selfplay (include/azb200_selfplay.h)."""
    _fields_ = [("game", C.c_int32), ("n", C.c_int32), ("max_plies", C.c_int32), ("pos_bytes", C.c_int32),
                ("temp_decay_moves", C.c_int32), ("temp_init", C.c_float), ("temp_endgame", C.c_float), ("forced_games", C.c_int32),
                ("seed", C.c_uint64), ("uid_stride", C.c_uint64), ("forced_uid0", C.c_uint64),
                ("states", C.c_void_p), ("steps", C.c_void_p), ("uids", C.c_void_p), ("st_pos", C.c_void_p),
                ("actions", C.c_void_p), ("fin_list", C.c_void_p), ("fin_count", C.c_void_p), ("forced", C.c_void_p),
                ("out_games", C.c_void_p), ("out_pos", C.c_void_p), ("out_counters", C.c_void_p),
                ("game_capacity", C.c_int32), ("reserved", C.c_int32)]


def pos_bytes(game: str) -> int:
    return int(_lib.lib().az_selfplay_pos_bytes(_G[game][0]))


def max_plies(game: str) -> int:
    return int(_lib.lib().az_selfplay_max_plies(_G[game][0]))


def shard_range(total: int, rank: int, world: int):
    """Contiguous game slice of rank r: [r*G/W, (r+1)*G/W)."""
    return (total * rank) // world, (total * (rank + 1)) // world


class Records:
    """Finished games as compact records: ``games`` uint8[m, 32] (az_sp_game) and ``pos`` uint8[P, pos_bytes] (az_sp_pos + policy
    target), game i owning rows [pos_start, pos_start + length) of ``pos``.  Lives on whatever device the tensors are on."""

    def __init__(self, game: str, games: torch.Tensor, pos: torch.Tensor):
        self.game, self.games, self.pos = game, games, pos

    def __len__(self):
        return int(self.games.shape[0])

    @property
    def positions(self) -> int:
        return int(self.pos.shape[0])

    # -- header fields (views; device of the tensors) ---------------------------------------------------------------
    def _i64(self):
        return self.games.view(torch.int64).reshape(-1, 4)

    def _i32(self):
        return self.games.view(torch.int32).reshape(-1, 8)

    @property
    def uid(self):
        return self._i64()[:, 0]

    @property
    def pos_start(self):
        return self._i64()[:, 1]

    @property
    def length(self):
        return self._i32()[:, 4]

    @property
    def winner(self):
        return self._i32()[:, 5]

    def cpu(self) -> "Records":
        return Records(self.game, self.games.cpu(), self.pos.cpu())

    def to(self, device) -> "Records":
        return Records(self.game, self.games.to(device), self.pos.to(device))

    def sorted_by_uid(self) -> "Records":
        """The same games ordered by uid, positions re-packed in that order (what a sharding-invariant comparison needs)."""
        m = len(self)
        if m == 0:
            return self
        order = torch.argsort(self.uid)
        length = self.length[order].long()
        new_start = torch.cumsum(length, 0) - length
        old_start = self.pos_start[order]
        rows = torch.repeat_interleave(old_start - new_start, length) + torch.arange(int(length.sum()), device=self.pos.device)
        games = self.games[order].clone()
        games.view(torch.int64).reshape(-1, 4)[:, 1] = new_start
        return Records(self.game, games, self.pos[rows])

    @staticmethod
    def cat(parts) -> "Records":
        """Concatenation (e.g. the per-rank segments of a gather): position rows are re-based."""
        parts = [p for p in parts if len(p)]
        if not parts:
            raise ValueError("no records")
        games, base = [], 0
        for p in parts:
            # a part may carry a position array with unused rows around its games: keep only what its games own
            lo = int(p.pos_start.min())
            hi = int((p.pos_start + p.length.long()).max())
            g = p.games.clone()
            g.view(torch.int64).reshape(-1, 4)[:, 1] += base - lo
            games.append((g, p.pos[lo:hi]))
            base += hi - lo
        return Records(parts[0].game, torch.cat([g for g, _ in games]), torch.cat([q for _, q in games]))

    # -- the reference's formats ----------------------------------------------------------------------------------
    def to_replay_tensors(self, td_steps: int = 0, stream: int | None = None) -> dict:
        """The 8 tensors of src/ReplayBuffer.py:12-19, one row per recorded position (terminal tuples included; row r belongs to
        position r of ``pos``), expanded by one kernel on the records' device (az_selfplay_expand_dev)."""
        if not self.games.is_cuda:
            raise RuntimeError("Records.to_replay_tensors runs on the GPU (no CPU fallback): move the records with .to('cuda')")
        gid, R, Cc, A = _G[self.game]
        P, dev = self.positions, self.games.device
        out = {
            "state": torch.empty((P, 3, R, Cc), dtype=torch.int8, device=dev),
            "prob": torch.empty((P, A), dtype=torch.float32, device=dev),
            "winner": torch.empty((P, 1), dtype=torch.int8, device=dev),
            "steps_to_end": torch.empty((P, 1), dtype=torch.int16, device=dev),
            "aux_target": torch.empty((P, 1), dtype=torch.int16, device=dev),
            "root_wdl": torch.empty((P, 3), dtype=torch.float32, device=dev),
            "valid_mask": torch.empty((P, A), dtype=torch.bool, device=dev),
            "future_root_wdl": torch.empty((P, 3), dtype=torch.float32, device=dev),
        }
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        games, pos = self.games.contiguous(), self.pos.contiguous()
        with torch.cuda.device(dev):
            rc = _lib.lib().az_selfplay_expand_dev(gid, len(self), games.data_ptr(), pos.data_ptr(), int(td_steps), out["state"].data_ptr(),
                                                   out["prob"].data_ptr(), out["winner"].data_ptr(), out["steps_to_end"].data_ptr(),
                                                   out["aux_target"].data_ptr(), out["root_wdl"].data_ptr(), out["valid_mask"].data_ptr(),
                                                   out["future_root_wdl"].data_ptr(), stream or None)
        if rc != 0:
            raise RuntimeError("az_selfplay_expand_dev failed (%d)" % rc)
        return out

    def unpack(self, td_steps: int = 0):
        """List of per-game dicts (in record order); ``tuples`` is what the reference's ``batch_self_play`` returns per game -
        (winner, tuple_of_training_tuples) with the element types of src/game.py:128-157 (numpy int32 scalars in the played
        positions, Python ints in the terminal tuple; ``future_root_wdl`` only when td_steps > 0)."""
        t = {k: v.cpu().numpy() for k, v in self.to_replay_tensors(td_steps).items()}
        uid, start, length, winner = (x.cpu().numpy() for x in (self.uid, self.pos_start, self.length, self.winner))
        games = []
        for i in range(len(self)):
            lo, L = int(start[i]), int(length[i])
            sl = slice(lo, lo + L)
            wz = t["winner"][sl, 0].astype(np.int32)
            ste = t["steps_to_end"][sl, 0].astype(np.int32)
            aux = t["aux_target"][sl, 0].astype(np.int32)
            rows = []
            for k in range(L):
                last = k == L - 1
                row = [t["state"][lo + k], t["prob"][lo + k], int(wz[k]) if last else wz[k], int(ste[k]) if last else ste[k],
                       int(aux[k]) if last else aux[k], t["root_wdl"][lo + k], t["valid_mask"][lo + k]]
                if td_steps > 0:
                    row.append(t["future_root_wdl"][lo + k])
                rows.append(tuple(row))
            games.append(dict(uid=int(uid[i]), winner=int(winner[i]), length=L, state=t["state"][sl], prob=t["prob"][sl],
                              root_wdl=t["root_wdl"][sl], future_root_wdl=t["future_root_wdl"][sl], winner_z=wz, steps_to_end=ste,
                              aux=aux, valid_mask=t["valid_mask"][sl], tuples=(int(winner[i]), tuple(rows))))
        return games


def unpack_records(rec: Records, game: str | None = None, td_steps: int = 0):
    return rec.unpack(td_steps)


def to_replay_tensors(rec: Records, game: str | None = None, td_steps: int = 0):
    return rec.to_replay_tensors(td_steps)


def save_replay_pt(path: str, tensors: dict):
    """Write the `.pt` layout that src/ReplayBuffer.py:25-62 saves/loads (same keys, order, dtypes; `_ptr`, `current_capacity`)."""
    n = tensors["state"].shape[0]
    order = ("state", "prob", "winner", "steps_to_end", "aux_target", "root_wdl", "valid_mask", "future_root_wdl")
    sd = {k: tensors[k].cpu() for k in order}
    sd["_ptr"], sd["current_capacity"] = n, n
    torch.save(sd, path)


def to_upload_payload(games) -> bytes:
    """The pickle an actor POSTs to the learner's /upload endpoint (client.py:367-373): {'__az__': True, 'data': [play_data, ...]}
    where play_data is the tuple of training tuples of one game (`Records.unpack(td)[i]['tuples'][1]`)."""
    import pickle
    return pickle.dumps({"__az__": True, "data": [g["tuples"][1] for g in games]}, protocol=pickle.HIGHEST_PROTOCOL)


class _Ring:
    """One output ring: game headers, position records, counters."""

    def __init__(self, game_capacity, pos_capacity, pb, device):
        self.games = torch.zeros((game_capacity, GAME_BYTES), dtype=torch.uint8, device=device)
        self.pos = torch.zeros((pos_capacity, pb), dtype=torch.uint8, device=device)
        self.counters = torch.zeros(4, dtype=torch.int64, device=device)     # games, positions, dropped, plies
        self.h_counters = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.ready = torch.cuda.Event()                                       # recorded when the ring was handed over
        self.free = None                                                      # event after which the ring may be refilled


class SelfPlay:
    def __init__(self, game, n_slots, n_playout, vl_batch, evaluator, search_cfg=None, temperature=1.0, temp_decay_moves=20,
                 temp_endgame=0.0, td_steps=10, seed=0, uid_base=0, uid_stride=None, device=None, out_capacity=None, cache_size=0,
                 forced_actions=None, forced_uid0=0):
        if not torch.cuda.is_available():
            raise RuntimeError("SelfPlay needs a CUDA device (no CPU fallback)")
        self.game = game
        self.gid, self.R, self.Cc, self.A = _G[game]
        self.T, self.pb = max_plies(game), pos_bytes(game)
        self.n, self.n_playout, self.K, self.td_steps = int(n_slots), int(n_playout), int(vl_batch), int(td_steps)
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
        d = dict(device=self.device)
        n, T, A = self.n, self.T, self.A
        self.out_capacity = int(out_capacity or 2 * n)
        self.states = torch.zeros((n, 32), dtype=torch.uint8, **d)
        self.steps = torch.zeros(n, dtype=torch.int32, **d)
        self.uids = (torch.arange(n, dtype=torch.int64, **d) + int(uid_base))
        self.st_pos = torch.zeros((n, T, self.pb), dtype=torch.uint8, **d)
        self.actions = torch.zeros(n, dtype=torch.int32, **d)
        self.fin_list = torch.zeros(n, dtype=torch.int32, **d)
        self.fin_count = torch.zeros(1, dtype=torch.int32, **d)
        self.rings = [_Ring(self.out_capacity, self.out_capacity * (T + 1), self.pb, self.device) for _ in range(2)]
        self.ring = 0
        self.counts = torch.zeros((n, A), dtype=torch.int32, **d)
        self.stats = torch.zeros((n, 6 + 8 * A), dtype=torch.float32, **d)
        self.forced = None
        if forced_actions is not None:
            fa = torch.as_tensor(np.ascontiguousarray(forced_actions), dtype=torch.int8)
            assert fa.ndim == 2 and fa.shape[1] == T, "forced_actions: int8[games, max_plies], -1 = play the searched move"
            self.forced = fa.to(self.device)
        self.buf = ds.LeafBuffers(n, n * max(self.K, 1), A, (self.R, self.Cc), self.device,
                                  unpacked=not isinstance(self.evaluator, ds.SyntheticEvaluator),
                                  planes=not isinstance(self.evaluator, ds.SyntheticEvaluator))
        self.buf.roots = self.states                     # the env states ARE the search roots
        self.sp = AzSelfplay(self.gid, n, T, self.pb, int(temp_decay_moves), float(temperature), float(temp_endgame),
                             0 if self.forced is None else int(self.forced.shape[0]), int(seed), int(uid_stride or n), int(forced_uid0),
                             self.states.data_ptr(), self.steps.data_ptr(), self.uids.data_ptr(), self.st_pos.data_ptr(),
                             self.actions.data_ptr(), self.fin_list.data_ptr(), self.fin_count.data_ptr(),
                             None if self.forced is None else self.forced.data_ptr(), None, None, None, self.out_capacity, 0)
        self._bind_ring()
        self._L = _lib.lib()
        rc = self._L.az_envs_reset_dev(self.gid, n, self.states.data_ptr(), self._stream())
        if rc != 0:
            raise RuntimeError("az_envs_reset_dev failed")
        self.plies = 0
        self.launches = 0

    def _bind_ring(self):
        r = self.rings[self.ring]
        self.sp.out_games, self.sp.out_pos, self.sp.out_counters = r.games.data_ptr(), r.pos.data_ptr(), r.counters.data_ptr()

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

    def finished(self) -> int:
        """Games in the current ring (one 32-byte D2H read; synchronises the current stream)."""
        return int(self.rings[self.ring].counters[0].item())

    def hand_over(self) -> _Ring:
        """Stream-ordered, no host synchronisation: the current ring is closed (its ``ready`` event marks the point on the current
        stream after which its contents are final, its counters are copied to pinned host memory) and the other ring - emptied -
        takes over.  The caller reads the ring on any stream after ``ready`` and sets ``ring.free`` to an event after its last use."""
        cur = torch.cuda.current_stream(self.device)
        r = self.rings[self.ring]
        r.h_counters.copy_(r.counters, non_blocking=True)
        r.ready.record(cur)
        self.ring ^= 1
        nxt = self.rings[self.ring]
        if nxt.free is not None:
            cur.wait_event(nxt.free)
            nxt.free = None
        nxt.counters.zero_()
        self._bind_ring()
        return r

    def drain(self) -> Records:
        """Everything finished since the last drain as ``Records`` (a snapshot: the ring is reused two drains later).  Raises if
        the ring overflowed - games were finished but not recorded - instead of returning a silently truncated set."""
        r = self.hand_over()
        r.ready.synchronize()
        m, p, dropped, _ = (int(x) for x in r.h_counters)
        if dropped:
            raise RuntimeError(f"self-play output ring overflowed: {dropped} finished games were dropped (capacity {self.out_capacity}); "
                               "drain more often or raise out_capacity")
        return Records(self.game, r.games[:m].clone(), r.pos[:p].clone())

    def run(self, target_games: int, max_plies: int | None = None) -> Records:
        """Play until `target_games` more games have finished (or `max_plies` plies were played); returns them all."""
        parts, done = [], 0
        start = self.plies
        while done < target_games and (max_plies is None or self.plies - start < max_plies):
            self.ply()
            cur = self.finished()                                  # the only host<->device traffic per ply
            if cur > self.out_capacity // 2:                       # keep the ring from filling up
                parts.append(self.drain())
                done += len(parts[-1])
            elif done + cur >= target_games:
                break
        rec = self.drain()
        if len(rec):
            parts.append(rec)
        if not parts:
            return Records(self.game, torch.zeros((0, GAME_BYTES), dtype=torch.uint8, device=self.device),
                           torch.zeros((0, self.pb), dtype=torch.uint8, device=self.device))
        return Records.cat(parts) if len(parts) > 1 else parts[0]

    @property
    def simulations(self):
        return self.plies * self.n * self.n_playout


def _all_gather_compact(game, games_buf, pos_buf, m, p, group, recv_games=None, recv_pos=None):
    """Counts first (16 bytes per rank), then exactly max(count) header rows and position rows per rank.  ``games_buf`` /
    ``pos_buf`` must hold at least max-over-ranks rows of memory (a ring does; exact-size records are padded by the caller).
    Returns (Records of all ranks in rank order, bytes received)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev, pb = games_buf.device, pos_buf.shape[1]
    cnt = torch.tensor([m, p], dtype=torch.int64, device=dev)
    cnts = torch.zeros((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts.view(-1), cnt, group=group)
    cn = cnts.cpu()
    mg, mp = int(cn[:, 0].max()), int(cn[:, 1].max())
    if mg == 0:
        return Records(game, games_buf[:0].clone(), pos_buf[:0].clone()), 0
    if games_buf.shape[0] < mg:
        games_buf = torch.cat([games_buf, games_buf.new_zeros((mg - games_buf.shape[0], GAME_BYTES))])
    if pos_buf.shape[0] < mp:
        pos_buf = torch.cat([pos_buf, pos_buf.new_zeros((mp - pos_buf.shape[0], pb))])
    rg = recv_games[:world * mg] if recv_games is not None else torch.empty((world * mg, GAME_BYTES), dtype=torch.uint8, device=dev)
    rp = (recv_pos[:world * mp * pb] if recv_pos is not None else torch.empty((world * mp * pb,), dtype=torch.uint8, device=dev)).view(world * mp, pb)
    dist.all_gather_into_tensor(rg, games_buf[:mg].contiguous(), group=group)
    dist.all_gather_into_tensor(rp, pos_buf[:mp].contiguous(), group=group)
    parts = [Records(game, rg[r * mg:r * mg + int(cn[r, 0])], rp[r * mp:r * mp + int(cn[r, 1])]) for r in range(world) if int(cn[r, 0])]
    return Records.cat(parts), world * (mg * GAME_BYTES + mp * pb)


def all_gather_records(rec: Records, group=None) -> Records:
    """Every rank contributes its finished games and receives everyone's, in rank order (blocking form of the exchange; works on
    whatever backend the process group has - NCCL for CUDA records, gloo for host records in the CPU tests)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rec
    return _all_gather_compact(rec.game, rec.games, rec.pos, len(rec), rec.positions, group)[0]


class TrajectoryExchange:
    """The one collective of the path: all-gather of the finished trajectories of every rank.

    Buffers are persistent (allocated once at ring capacity); a call first gathers the per-rank (games, positions) counts - 16
    bytes per rank - and then exactly ``max(count)`` rows of headers and of position records per rank straight out of the ring, so
    the bytes on the wire follow the payload, not the ring capacity.  Everything is issued on the exchange's own stream:
    ``submit(ring)`` orders it after the ring's ``ready`` event, so the gather of batch i runs while the main stream searches batch
    i + 1 (call it after batch i + 1 was enqueued: reading the counts blocks the host until batch i is complete);
    ``collect()`` returns the gathered ``Records`` of the oldest submitted batch."""

    def __init__(self, game: str, game_capacity: int, device, group=None):
        import torch.distributed as dist
        self.game, self.device, self.group = game, torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pb, T = pos_bytes(game), max_plies(game)
        self.gcap, self.pcap = int(game_capacity), int(game_capacity) * (T + 1)
        self.stream = torch.cuda.Stream(device=self.device)
        self.recv_games = self.recv_pos = None
        if self.world > 1:
            self.recv_games = torch.empty((self.world * self.gcap, GAME_BYTES), dtype=torch.uint8, device=self.device)
            self.recv_pos = torch.empty((self.world * self.pcap * self.pb,), dtype=torch.uint8, device=self.device)
        self.pending = []
        self.bytes_gathered = 0
        self.ms = []

    def submit(self, ring: _Ring):
        st = self.stream
        st.wait_event(ring.ready)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ring.ready.synchronize()                                     # the ring's counters are in pinned host memory now
        m, p, dropped = int(ring.h_counters[0]), int(ring.h_counters[1]), int(ring.h_counters[2])
        if dropped:
            raise RuntimeError(f"self-play output ring overflowed before the exchange ({dropped} games dropped)")
        with torch.cuda.stream(st):
            t0.record(st)
            if self.world == 1:
                rec = Records(self.game, ring.games[:m].clone(), ring.pos[:p].clone())
            else:
                rec, nbytes = _all_gather_compact(self.game, ring.games, ring.pos, m, p, self.group, self.recv_games, self.recv_pos)
                self.bytes_gathered += nbytes
            t1.record(st)
            free = torch.cuda.Event()
            free.record(st)
        ring.free = free
        self.pending.append((rec, t0, t1))

    def collect(self) -> Records:
        rec, t0, t1 = self.pending.pop(0)
        t1.synchronize()
        self.ms.append(t0.elapsed_time(t1))
        return rec

    def gather(self, ring: _Ring) -> Records:
        self.submit(ring)
        return self.collect()
