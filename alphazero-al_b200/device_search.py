"""Device-resident playout loop: the reference wrapper's per-move loop (src/MCTS_cpp.py:89-359 - one non-VL
warm-up simulation, then ceil((n-1)/K) virtual-loss iterations with cur_K = min(K, remaining)) with every buffer
in HBM and no host synchronisation.  select -> evaluator -> backprop are three stream-ordered launches per
iteration; the evaluator is either the synthetic CUDA evaluator (csrc/az_eval.cu) or any callable working on the
torch tensors in ``LeafBuffers`` (e.g. a CNN consuming ``planes``).

PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

SYN_MODES = {"hash": 0, "equivariant": 1, "constant": 2}


class LeafBuffers:
    """Device buffers of one playout loop: az_root[n] in, az_leaf[rows] out of select, evaluation tuple into backprop.
    `unpacked=True` additionally allocates the reference-format arrays (int8 boards, masks, ...) and `planes=True`
    the CNN input tensor, both filled on demand by `unpack()`."""

    def __init__(self, n: int, rows: int, A: int, board_shape, device, unpacked: bool = False, planes: bool = False):
        dv = dict(device=device)
        self.n, self.rows, self.A, self.board_shape = n, rows, A, tuple(board_shape)
        self.gid = 0 if tuple(board_shape) == (6, 7) else 1
        self.roots = torch.zeros((n, 32), dtype=torch.uint8, **dv)
        self.leaves = torch.zeros((rows, 32), dtype=torch.uint8, **dv)
        self.policy = torch.empty((rows, A), dtype=torch.float32, **dv)
        self.d, self.p1w, self.p2w, self.ml = (torch.empty(rows, dtype=torch.float32, **dv) for _ in range(4))
        self.boards = self.td = self.tp1 = self.tp2 = self.is_term = self.turns = self.sym = self.mask = self.planes = None
        if unpacked:
            self.boards = torch.empty((rows, *board_shape), dtype=torch.int8, **dv)
            self.td, self.tp1, self.tp2 = (torch.empty(rows, dtype=torch.float32, **dv) for _ in range(3))
            self.is_term = torch.empty(rows, dtype=torch.uint8, **dv)
            self.turns = torch.empty(rows, dtype=torch.int32, **dv)
            self.sym = torch.empty(rows, dtype=torch.int32, **dv)
            self.mask = torch.empty((rows, A), dtype=torch.uint8, **dv)
        if planes:
            self.planes = torch.empty((rows, 3, *board_shape), dtype=torch.float32, **dv)

    def pack_roots(self, boards: torch.Tensor, turns: torch.Tensor, stream: int):
        """int8[n,R,C] boards + int32[n] turns (CUDA tensors) -> az_root[n]."""
        rc = _lib.lib().az_pack_roots_dev(self.gid, self.n, boards.data_ptr(), turns.data_ptr(), self.roots.data_ptr(),
                                          stream or None)
        if rc != 0:
            raise RuntimeError("az_pack_roots_dev failed (%d)" % rc)

    def unpack(self, rows: int, stream: int, row0: int = 0):
        """az_leaf[row0 : row0 + rows] -> whichever reference-format arrays / CNN planes were allocated."""
        p = lambda t: t[row0:].data_ptr() if t is not None else None
        rc = _lib.lib().az_unpack_leaves_dev(self.gid, rows, p(self.leaves), p(self.boards), p(self.td), p(self.tp1),
                                             p(self.tp2), p(self.is_term), p(self.turns), p(self.sym), p(self.mask),
                                             p(self.planes), stream or None)
        if rc != 0:
            raise RuntimeError("az_unpack_leaves_dev failed (%d)" % rc)


class SyntheticEvaluator:
    """Device twin of evaluators.HashEvaluator (one kernel launch on the az_leaf records, no host round trip)."""

    def __init__(self, game: str, mode: str = "hash"):
        self.gid = {"Connect4": 0, "Othello": 1}[game]
        self.mode = SYN_MODES[mode]
        self.launches = 0

    def __call__(self, buf: LeafBuffers, rows: int, stream: int, row0: int = 0):
        rc = _lib.lib().az_eval_synthetic_dev(
            self.gid, self.mode, rows, buf.leaves[row0:].data_ptr(), buf.policy[row0:].data_ptr(), buf.d[row0:].data_ptr(),
            buf.p1w[row0:].data_ptr(), buf.p2w[row0:].data_ptr(), buf.ml[row0:].data_ptr(), stream or None)
        if rc != 0:
            raise RuntimeError("az_eval_synthetic_dev failed (%d)" % rc)
        self.launches += 1


class NetEvaluator:
    """Fused evaluator boundary: az_leaf records -> CNN input planes + legal masks (one unpack
    kernel) -> ``net.predict_device(planes, mask)`` -> backprop tuple (one finalize kernel).  No host copy, no sync.
    `buf` must have been created with unpacked=True, planes=True.

    `graph_rows`: batches of at most this many rows replay the network from a CUDA graph captured per (buffer, row range):
    a small batch (the reference's own 100-game configuration evaluates 100-400 leaves per iteration) is bound by the ~150
    kernel launches of an eager forward pass, not by the GPU.  Larger batches run eagerly (nothing to gain, and a graph keeps
    its activations allocated).  A network that cannot be captured falls back to eager calls for good."""

    def __init__(self, net, graph_rows: int = 8192):
        self.net = net
        self._wdl = self._aux = None
        self.graph_rows = int(graph_rows)
        self._graphs = {}
        self._fp, self._fp_params, self._fp_calls, self._fp_gen = None, None, 0, 0
        self.graph_replays = 0

    def _check_weights(self):
        """Captured graphs read the parameter tensors that existed at capture time: in-place updates (load_state_dict) are seen,
        re-created parameters (net.to(dtype), net.half(), a swapped module) are not - drop the graphs when any storage moved.
        The fingerprint covers every parameter and buffer (data pointer + dtype); modules that re-create tensors do it through
        ``_apply``, which is hooked once so the walk (~90 us) only happens after such a call or every 64th evaluation."""
        self._fp_calls += 1
        net = self.net
        inner = getattr(net, "net", net)                     # ReferenceNetAdapter wraps the module
        if self._fp_params is None and isinstance(inner, torch.nn.Module) and not getattr(inner, "_azb200_apply_hooked", False):
            orig_apply = inner._apply

            def hooked(fn, *a, _orig=orig_apply, _mod=inner, **k):
                _mod._azb200_generation = getattr(_mod, "_azb200_generation", 0) + 1
                return _orig(fn, *a, **k)
            try:
                inner._apply = hooked
                inner._azb200_apply_hooked = True
            except Exception:
                pass
        gen = getattr(inner, "_azb200_generation", 0)
        if self._fp_params is None or gen != self._fp_gen or self._fp_calls % 64 == 0:
            self._fp_gen = gen
            ps = []
            for name in ("parameters", "buffers"):
                it = getattr(net, name, None)
                if callable(it):
                    ps.extend(it())
            self._fp_params = ps
            fp = tuple((p.data_ptr(), p.dtype) for p in ps if hasattr(p, "data_ptr"))
            if fp != self._fp:
                self._graphs.clear()
                self._fp = fp

    def _forward(self, buf, r, rows):
        probs, wdl_rel, aux = self.net.predict_device(buf.planes[r], buf.mask[r])
        buf.policy[r].copy_(probs.reshape(rows, buf.A))
        self._wdl[r].copy_(wdl_rel.reshape(rows, 3))
        self._aux[r].copy_(aux.reshape(rows))

    def _captured(self, buf, r, rows):
        """Graph of _forward for this row range (inputs and outputs are slices of persistent buffers), or None."""
        key = (buf.planes.data_ptr(), buf.policy.data_ptr(), self._wdl.data_ptr(), r.start, rows)
        g = self._graphs.get(key)
        if g is None and self.graph_rows > 0:
            cur = torch.cuda.current_stream(buf.leaves.device)
            try:
                side = torch.cuda.Stream(device=buf.leaves.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side), torch.no_grad():       # warm-up off the capture: cuDNN plans, lazy initialisation
                    for _ in range(2):
                        self._forward(buf, r, rows)
                cur.wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g), torch.no_grad():
                    self._forward(buf, r, rows)
            except Exception:                                        # e.g. a predict_device that synchronises: stay eager
                g = False
                self.graph_rows = 0
                torch.cuda.synchronize(buf.leaves.device)
            self._graphs[key] = g
        return g or None

    def __call__(self, buf: LeafBuffers, rows: int, stream: int, row0: int = 0):
        if self._wdl is None or self._wdl.shape[0] < buf.rows:
            self._wdl = torch.empty((buf.rows, 3), dtype=torch.float32, device=buf.leaves.device)
            self._aux = torch.empty(buf.rows, dtype=torch.float32, device=buf.leaves.device)
            self._graphs.clear()
        r = slice(row0, row0 + rows)
        buf.unpack(rows, stream, row0)
        g = None
        if rows <= self.graph_rows:
            self._check_weights()
            g = self._captured(buf, r, rows)
        if g is not None:
            g.replay()
            self.graph_replays += 1
        else:
            self._forward(buf, r, rows)
        rc = _lib.lib().az_eval_finalize_dev(rows, buf.leaves[row0:].data_ptr(), self._wdl[row0:].data_ptr(), self._aux[row0:].data_ptr(),
                                             buf.d[row0:].data_ptr(), buf.p1w[row0:].data_ptr(), buf.p2w[row0:].data_ptr(),
                                             buf.ml[row0:].data_ptr(), stream or None)
        if rc != 0:
            raise RuntimeError("az_eval_finalize_dev failed (%d)" % rc)


class ReferenceNetAdapter:
    """Device contract for an UNMODIFIED network of the reference (src/environments/{Connect4,Othello}/Network.py::CNN): the
    post-processing of its ``predict`` (Connect4/Network.py:267-288, Othello/Network.py:235-261 - bf16 autocast off the CPU,
    exp of the log-policy and log-WDL, aux = steps_norm * aux_target_offset for Connect4, atan(disc_diff / score_scale) * 2/pi
    for Othello) on tensors that stay where they are: no pinned staging, no ``.cpu()``, no ``torch.cuda.synchronize()``.
    ``BatchedMCTS.batch_playout`` wraps such a network by itself, so ``src/player.py`` reaches the device-resident loop without
    any change."""

    def __init__(self, net, game: str):
        self.net, self.game = net, game

    @staticmethod
    def accepts(net) -> bool:
        """A torch module with the reference's forward contract whose weights live on a CUDA device."""
        if not isinstance(net, torch.nn.Module) or not hasattr(net, "aux_target_offset") or hasattr(net, "predict_device"):
            return False
        p = next(net.parameters(), None)
        return p is not None and p.is_cuda

    def parameters(self):
        return self.net.parameters()

    def buffers(self):
        return self.net.buffers()

    @property
    def score_scale(self):
        return getattr(self.net, "score_scale", 8.0)

    @score_scale.setter
    def score_scale(self, v):
        self.net.score_scale = v

    @torch.no_grad()
    def predict_device(self, planes, action_mask):
        kind = planes.device.type
        with torch.autocast(kind, dtype=torch.bfloat16, enabled=kind != "cpu"):
            log_prob, value_log_prob, aux = self.net(planes, action_mask=action_mask.bool())
        # the dtype flow of the reference's predict(): policy .float().exp(); WDL exp() in the head's dtype then .float(); the aux
        # scaling (and Othello's atan) in the head's dtype - bf16 under autocast - and .float() last
        probs, wdl = log_prob.float().exp(), value_log_prob.exp().float()
        aux = aux * float(self.net.aux_target_offset)
        if self.game == "Othello":
            import math
            aux = torch.atan(aux / self.score_scale) * (2.0 / math.pi)
        return probs, wdl, aux.float().reshape(-1)


class EvalCache:
    """Device evaluation cache (include/azb200_cache.h; reference: src/Cache.py LRUCache used by src/MCTS_cpp.py)."""

    def __init__(self, game: str, cache_size: int, device: int | None = None):
        self.gid = {"Connect4": 0, "Othello": 1}[game]
        log2 = max(4, int(2 * max(cache_size, 1) - 1).bit_length())       # >= 2x the requested entry count
        dev = torch.cuda.current_device() if device is None else int(device)
        self._L = _lib.lib()
        self._c = self._L.az_evalcache_create(self.gid, log2, dev)
        if not self._c:
            raise RuntimeError("az_evalcache_create failed")
        self._c = C.c_void_p(self._c)

    def __del__(self):
        c, self._c = getattr(self, "_c", None), None
        if c:
            self._L.az_evalcache_destroy(c)

    def clear(self, stream=0):
        """refresh_cache / score_scale invalidation: cached values are stale after a weight reload."""
        if self._L.az_evalcache_clear_dev(self._c, stream or None) != 0:
            raise RuntimeError("az_evalcache_clear_dev failed")

    def stats(self):
        import numpy as np
        out = np.zeros(4, np.uint64)
        if self._L.az_evalcache_stats(self._c, out.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("az_evalcache_stats failed")
        dups = np.zeros(1, np.uint64)
        if self._L.az_evalcache_dups(self._c, dups.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("az_evalcache_dups failed")
        return dict(lookups=int(out[0]), hits=int(out[1]), inserts=int(out[2]), capacity=int(out[3]), dups=int(dups[0]))


class CachedNetEvaluator(NetEvaluator):
    """NetEvaluator that probes the device cache first; only misses reach the network, and of the missing leaves that hold
    the same position only one does (`dedup`; self-play batches repeat openings in every slot).  One 4-byte D2H read per
    iteration (the miss count) is the price of a dynamically sized network batch."""

    def __init__(self, net, cache: EvalCache, dedup: bool = True, graph_rows: int = 8192):
        super().__init__(net, graph_rows=graph_rows)    # graphs are captured per batch-size bucket here (see _net_rows)
        self.cache = cache
        self.dedup = dedup
        self._miss_idx = self._miss_cnt = self._dup_of = self._stage = None
        self.net_rows = 0

    shardable = False       # one miss counter / index buffer per evaluator: drive it on the whole batch

    def _eager_rows(self, planes, mask, mb, A):
        probs, wdl_rel, aux = self.net.predict_device(planes, mask)
        return probs.reshape(mb, A).float().contiguous(), wdl_rel.reshape(mb, 3).float().contiguous(), aux.reshape(mb).float().contiguous()

    def _net_rows(self, buf, idx, mb):
        """Network outputs (probs[mb,A], wdl_rel[mb,3], aux[mb]) of the leaves `idx`.  Small buckets gather the rows into static
        staging tensors and replay the forward pass from a CUDA graph captured per bucket size (see NetEvaluator)."""
        if mb <= self.graph_rows:
            self._check_weights()
            if self._stage is None or self._stage[0].shape[0] < min(self.graph_rows, buf.rows) or self._stage[2] != buf.planes.data_ptr():
                cap = min(self.graph_rows, buf.rows)
                self._stage = (torch.zeros((cap, *buf.planes.shape[1:]), dtype=buf.planes.dtype, device=buf.planes.device),
                               torch.ones((cap, buf.A), dtype=buf.mask.dtype, device=buf.planes.device), buf.planes.data_ptr())
                self._graphs.clear()
            sp, sm = self._stage[0][:mb], self._stage[1][:mb]
            g = self._graphs.get(mb)
            if g is None:
                cur = torch.cuda.current_stream(buf.leaves.device)
                try:
                    side = torch.cuda.Stream(device=buf.leaves.device)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side), torch.no_grad():
                        for _ in range(2):
                            self._eager_rows(sp, sm, mb, buf.A)
                    cur.wait_stream(side)
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg), torch.no_grad():
                        outs = self._eager_rows(sp, sm, mb, buf.A)
                    g = (cg, outs)
                except Exception:
                    g = False
                    self.graph_rows = 0
                    torch.cuda.synchronize(buf.leaves.device)
                self._graphs[mb] = g
            if g:
                torch.index_select(buf.planes, 0, idx, out=sp)
                torch.index_select(buf.mask, 0, idx, out=sm)
                g[0].replay()
                self.graph_replays += 1
                return g[1]
        return self._eager_rows(buf.planes.index_select(0, idx), buf.mask.index_select(0, idx), mb, buf.A)

    def __call__(self, buf: LeafBuffers, rows: int, stream: int, row0: int = 0):
        assert row0 == 0, "CachedNetEvaluator evaluates the whole batch at once"
        L, dev = _lib.lib(), buf.leaves.device
        if self._wdl is None or self._wdl.shape[0] < buf.rows:
            self._wdl = torch.empty((buf.rows, 3), dtype=torch.float32, device=dev)
            self._aux = torch.empty(buf.rows, dtype=torch.float32, device=dev)
            self._miss_idx = torch.zeros(buf.rows, dtype=torch.int32, device=dev)
            self._miss_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            self._dup_of = torch.empty(buf.rows, dtype=torch.int32, device=dev)
        self._miss_cnt.zero_()
        if self.dedup:
            rc = L.az_evalcache_lookup_dedup_dev(self.cache._c, rows, buf.leaves.data_ptr(), buf.policy.data_ptr(), self._wdl.data_ptr(),
                                                 self._aux.data_ptr(), self._miss_idx.data_ptr(), self._miss_cnt.data_ptr(),
                                                 self._dup_of.data_ptr(), stream or None)
        else:
            rc = L.az_evalcache_lookup_dev(self.cache._c, rows, buf.leaves.data_ptr(), buf.policy.data_ptr(), self._wdl.data_ptr(),
                                           self._aux.data_ptr(), self._miss_idx.data_ptr(), self._miss_cnt.data_ptr(), stream or None)
        if rc != 0:
            raise RuntimeError("az_evalcache_lookup_dev failed")
        m = int(self._miss_cnt.item())
        if m > 0:
            buf.unpack(rows, stream)
            # the network sees a few batch sizes only (every new shape costs cuDNN / SDPA plan selection: tens of ms): the index
            # list is read up to the next bucket - the entries past m are older row numbers, their outputs are dropped
            mb = _bucket(m, buf.rows)
            pm, wm, am = self._net_rows(buf, self._miss_idx[:mb].long(), mb)
            rc = L.az_evalcache_insert_dev(self.cache._c, m, buf.leaves.data_ptr(), self._miss_idx.data_ptr(), pm.data_ptr(), wm.data_ptr(),
                                           am.data_ptr(), buf.policy.data_ptr(), self._wdl.data_ptr(), self._aux.data_ptr(), stream or None)
            if rc != 0:
                raise RuntimeError("az_evalcache_insert_dev failed")
            self.net_rows += m
            if self.dedup:                  # duplicates copy the row of the leaf that was evaluated for them
                rc = L.az_evalcache_resolve_dups_dev(self.cache._c, rows, self._dup_of.data_ptr(), buf.policy.data_ptr(),
                                                     self._wdl.data_ptr(), self._aux.data_ptr(), stream or None)
                if rc != 0:
                    raise RuntimeError("az_evalcache_resolve_dups_dev failed")
        rc = L.az_eval_finalize_dev(rows, buf.leaves.data_ptr(), self._wdl.data_ptr(), self._aux.data_ptr(), buf.d.data_ptr(),
                                    buf.p1w.data_ptr(), buf.p2w.data_ptr(), buf.ml.data_ptr(), stream or None)
        if rc != 0:
            raise RuntimeError("az_eval_finalize_dev failed (%d)" % rc)


def _bucket(m: int, cap: int, floor: int = 256) -> int:
    """Smallest batch size of the form {4,5,6,7} * 2^k >= max(m, floor), at most `cap` (<= 25 % padding, 4 shapes per octave)."""
    m = max(m, min(floor, cap))
    k = max(m.bit_length() - 3, 0)
    b = -(-m // (1 << k)) << k
    return min(b, cap)


_SHARD_STREAMS = {}


def auto_shards(n: int) -> int:
    """Independent tree shards driven on their own streams: 8192 trees per shard, at most 8 (measured on B200, Connect4
    n=200 K=4, 65 536 trees, loop replayed from its CUDA graph: 1 shard 2.25, 4 shards 2.80, 8 shards 3.31, 16 shards 3.30 G
    simulations/s).  Small batches stay whole: they are latency bound."""
    import os
    if os.environ.get("AZB200_SHARDS"):
        return max(1, int(os.environ["AZB200_SHARDS"]))
    return max(1, min(8, n // 8192))


def iteration_schedule(n_playout: int, K: int):
    """K of every iteration of the wrapper's loop (src/MCTS_cpp.py:217-357): 0 = the non-VL warm-up simulation, then virtual-loss
    batches of K (the last one smaller); K <= 1: n_playout non-VL simulations."""
    if K <= 1:
        return [0] * n_playout
    if n_playout <= 0:
        return []
    rem = n_playout - 1
    return [0] + [K] * (rem // K) + ([rem % K] if rem % K else [])


def playout_device(engine, buf: LeafBuffers, n_playout: int, K: int, evaluator, stream: int | None = None,
                   on_select=None, shards: int | None = None, iters=None):
    """Run `n_playout` simulations per tree entirely on the device from the roots in `buf.roots` (see
    LeafBuffers.pack_roots).  `on_select(rows, fn)` (optional) wraps each select launch (bench.py times the dominant
    kernel with CUDA events through it).  Returns the number of kernels launched.

    `iters` (optional) = an explicit list of per-iteration K values instead of the whole schedule of (n_playout, K).

    `shards` > 1 splits the batch into that many independent tree ranges, each running its own
    select -> evaluate -> backprop chain on its own stream, so that the latency-bound tree kernels of one shard overlap
    the evaluation of another.  Trees are independent and the RNG keys do not depend on the split, so the result is
    identical to the unsharded loop (tests/test_gpu_mcts.py).  None = auto_shards(n)."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    n = engine.get_num_envs()
    if shards is None:
        shards = auto_shards(n)
    if not getattr(evaluator, "shardable", True) or n < 64 * shards:
        shards = 1
    explicit = iters is not None            # a slice of the schedule (time-budgeted searches issue it in chunks)
    if not explicit:
        iters = iteration_schedule(n_playout, K)
    if isinstance(evaluator, SyntheticEvaluator) and on_select is None and not explicit:
        # everything is inside the library: let it drive the loop (no per-launch Python / ctypes cost)
        assert n * max(K, 1) <= buf.rows
        return engine.playout_synthetic_dev(evaluator.mode, n_playout, K, shards, buf.roots.data_ptr(), buf.leaves.data_ptr(),
                                            buf.policy.data_ptr(), buf.d.data_ptr(), buf.p1w.data_ptr(), buf.p2w.data_ptr(),
                                            buf.ml.data_ptr(), stream)
    if shards > 1:
        return _playout_sharded(engine, buf, iters, evaluator, stream, shards, on_select)
    launches = 0
    # torch evaluators launch on torch's current stream: make that the stream the engine kernels run on
    import contextlib
    ext = None
    if buf.leaves.is_cuda:
        ext = torch.cuda.ExternalStream(stream, device=buf.leaves.device) if stream else torch.cuda.default_stream(buf.leaves.device)
    for k in iters:
        rows = n * max(k, 1)
        assert rows <= buf.rows
        if on_select is not None:
            on_select(rows, lambda: engine.search_dev(k, buf.roots.data_ptr(), buf.leaves.data_ptr(), stream))
        else:
            engine.search_dev(k, buf.roots.data_ptr(), buf.leaves.data_ptr(), stream)
        with (torch.cuda.stream(ext) if ext is not None else contextlib.nullcontext()):
            evaluator(buf, rows, stream)
        # is_term / sym ids: the engine uses what it remembered from the matching search
        engine.backprop_dev(k, buf.policy.data_ptr(), buf.d.data_ptr(), buf.p1w.data_ptr(), buf.p2w.data_ptr(),
                            buf.ml.data_ptr(), 0, 0, stream)
        launches += 3
    return launches


def _playout_sharded(engine, buf, iters, evaluator, stream, shards, on_select=None):
    n = engine.get_num_envs()
    dev = buf.leaves.device
    key = (dev.index, shards)
    if key not in _SHARD_STREAMS:
        _SHARD_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(shards)]
    side = _SHARD_STREAMS[key]
    per = ((n + shards - 1) // shards + 31) // 32 * 32            # shard starts are multiples of 32
    ranges = [(lo, min(per, n - lo)) for lo in range(0, n, per)]
    main = torch.cuda.ExternalStream(stream, device=dev)
    engine.stream_handover_dev(stream)                           # main stream now follows everything queued so far
    fork = torch.cuda.Event()
    fork.record(main)
    for st in side[:len(ranges)]:
        st.wait_event(fork)
    ptrs = (buf.policy.data_ptr(), buf.d.data_ptr(), buf.p1w.data_ptr(), buf.p2w.data_ptr(), buf.ml.data_ptr())
    launches = 0
    # every shard owns rows [lo * kmax, (lo + cnt) * kmax) for the whole loop: shards run ahead of each other, and with the
    # whole-batch layout (row = tree * K + k) one shard's K = 4 rows would overlap another's K = 1 / remainder rows
    kmax = max([1] + iters)
    assert n * kmax <= buf.rows
    for it, k in enumerate(iters):
        kk = max(k, 1)
        for j, (lo, cnt) in enumerate(ranges):
            st = side[j]
            cs = st.cuda_stream
            row0 = lo * kmax
            if on_select is not None:
                on_select(cnt * kk, lambda: engine.search_range_dev(k, buf.roots.data_ptr(), buf.leaves.data_ptr(), lo, cnt, row0, j == 0, cs), st)
            else:
                engine.search_range_dev(k, buf.roots.data_ptr(), buf.leaves.data_ptr(), lo, cnt, row0, j == 0, cs)
            with torch.cuda.stream(st):                          # torch evaluators launch on the current stream
                evaluator(buf, cnt * kk, cs, row0)
            engine.backprop_range_dev(k, *ptrs, lo, cnt, row0, 0, 0, cs)
            launches += 3
    for st in side[:len(ranges)]:
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
    engine.stream_handover_dev(stream)
    return launches
