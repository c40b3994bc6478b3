"""Host-side mirror of the reference search wrapper ``src/MCTS_cpp.py::BatchedMCTS`` (same constructor, same
methods, same return types) on top of the CUDA engine.

Two evaluation paths:

* device-resident (B200 path): when ``pv_func`` is a ``device_search.SyntheticEvaluator`` or exposes
  ``predict_device(planes f32[B,3,R,C], mask u8[B,A]) -> (probs[B,A], wdl_rel[B,3], aux[B])`` on CUDA tensors, the
  whole per-move loop of src/MCTS_cpp.py:89-359 runs as stream-ordered kernel launches - leaves are encoded
  straight into the network's input tensor, no ``.cpu()``, no ``synchronize`` until visit counts are read.
* host predict (drop-in path): any object with the reference's ``predict(state_np, action_mask=...)`` contract is
  driven exactly like the reference wrapper does, through the host-buffer entry points of ``mcts_cpp``,
  including the LRU evaluation cache of src/Cache.py (key = symmetrised leaf board bytes + turn byte).
"""
from __future__ import annotations

import time
from collections import OrderedDict

import numpy as np

from . import _lib, mcts_cpp

_BACKENDS = {"Connect4": mcts_cpp.BatchedMCTS_Connect4, "Othello": mcts_cpp.BatchedMCTS_Othello}


def _default_convert_board(board, turns):
    """3-plane relative encoding (src/MCTS_cpp.py:15-20)."""
    plane_x = (board == turns[:, None, None]).astype(np.float32)
    plane_o = (board == -turns[:, None, None]).astype(np.float32)
    plane_turn = np.ones_like(board, dtype=np.float32) * turns[:, None, None]
    return np.stack([plane_x, plane_o, plane_turn], axis=1)


def _relative_wdl_to_absolute(wdl_rel, turns):
    d, w, l = wdl_rel[:, 0], wdl_rel[:, 1], wdl_rel[:, 2]
    return d, np.where(turns == 1, w, l), np.where(turns == 1, l, w)


class _LRU:
    """Semantics of src/Cache.py:5-58 (most-recent at the front, evict from the back)."""

    def __init__(self, cap):
        self._cap, self._od = cap, OrderedDict()

    def __contains__(self, k):
        return k in self._od

    def __len__(self):
        return len(self._od)

    def get(self, k):
        self._od.move_to_end(k, last=False)
        return self._od[k]["value"]

    def put(self, k, v):
        if k in self._od:
            self._od[k]["value"] = v
        else:
            self._od[k] = {"state": k, "value": v}
        self._od.move_to_end(k, last=False)
        if len(self._od) > self._cap:
            self._od.popitem(last=True)


class BatchedMCTS:
    def __init__(self, batch_size, c_init, c_base, alpha, n_playout, game_name="Connect4", board_converter=None,
                 cache_size=0, noise_epsilon=0.25, fpu_reduction=0.4, use_symmetry=True, mlh_slope=0.0, mlh_cap=0.2,
                 value_decay=1.0, score_utility_factor=0.0, score_scale=8.0, device=None):
        backend_cls = _BACKENDS[game_name]
        self.mcts = backend_cls(batch_size, device=device)
        cfg = self.mcts.config
        cfg.c_init, cfg.c_base, cfg.dirichlet_alpha = c_init, c_base, alpha
        cfg.noise_epsilon, cfg.fpu_reduction, cfg.use_symmetry = noise_epsilon, fpu_reduction, use_symmetry
        cfg.mlh_slope, cfg.mlh_cap = mlh_slope, mlh_cap
        cfg.score_utility_factor, cfg.score_scale, cfg.value_decay = score_utility_factor, score_scale, value_decay
        self.n_playout, self.batch_size = n_playout, batch_size
        self.action_size, self.board_shape = backend_cls.action_size, backend_cls.board_shape
        self._convert_board = board_converter or _default_convert_board
        self._custom_converter = board_converter is not None      # the device path encodes leaves itself (default planes only)
        self.cache = _LRU(cache_size) if cache_size > 0 else None      # host LRU, used by the numpy `predict` path only
        self._cache_size, self._dev_cache = cache_size, None
        self._evaluators = {}
        self._adapters = {}
        self._game_name = game_name
        self._rollout_eval = None
        self._dev = None            # lazily created device-side state (torch tensors)

    # ------------------------------------------------------------------------------------------------------
    # device-resident path
    # ------------------------------------------------------------------------------------------------------
    def _device_state(self, K):
        import torch
        from . import device_search as ds
        rows = self.batch_size * max(K, 1)
        if self._dev is None or self._dev["buf"].rows < rows:
            dev = torch.device("cuda", self.mcts._device if hasattr(self.mcts, "_device") else torch.cuda.current_device())
            buf = ds.LeafBuffers(self.batch_size, rows, self.action_size, self.board_shape, dev, unpacked=True, planes=True)
            n = self.batch_size
            self._dev = dict(
                buf=buf, dev=dev,
                boards=torch.empty((n, *self.board_shape), dtype=torch.int8, device=dev),
                turns=torch.empty(n, dtype=torch.int32, device=dev),
                h_boards=torch.empty((n, *self.board_shape), dtype=torch.int8).pin_memory(),
                h_turns=torch.empty(n, dtype=torch.int32).pin_memory(),
                counts=torch.empty((n, self.action_size), dtype=torch.int32, device=dev),
                h_counts=torch.empty((n, self.action_size), dtype=torch.int32).pin_memory(),
                wdl=torch.empty((rows, 3), dtype=torch.float32, device=dev),
                aux=torch.empty(rows, dtype=torch.float32, device=dev))
            self._dev["h_boards_np"] = self._dev["h_boards"].numpy()
            self._dev["h_turns_np"] = self._dev["h_turns"].numpy()
        return self._dev

    def _playout_device(self, pv_func, current_boards, turns, max_n, K, time_budget=None):
        import torch
        from . import device_search as ds
        if isinstance(pv_func, ds.SyntheticEvaluator) and time_budget is None:
            # everything is inside the library: host arrays in, the loop pipelined shard by shard, visit counts left in pinned memory
            # for the get_visits_count() that follows (src/player.py:333-343)
            self._last_evaluator = pv_func
            self.mcts.playout_synthetic_host(pv_func.mode, max_n, K, ds.auto_shards(self.batch_size), current_boards, turns)
            return
        st = self._device_state(K)
        buf = st["buf"]
        stream = torch.cuda.current_stream().cuda_stream
        # one pass from the caller's arrays (any dtype: Env.board is float32, src/MCTS_cpp.py:101-102) into pinned staging
        np.copyto(st["h_boards_np"], current_boards, casting="unsafe")
        np.copyto(st["h_turns_np"], turns, casting="unsafe")
        st["boards"].copy_(st["h_boards"], non_blocking=True)
        st["turns"].copy_(st["h_turns"], non_blocking=True)
        buf.pack_roots(st["boards"], st["turns"], stream)
        if isinstance(pv_func, ds.SyntheticEvaluator):
            evaluator = pv_func
        else:                               # one evaluator object per network: it owns the captured CUDA graphs of the forward pass
            key = (id(pv_func), self._cache_size > 0)
            evaluator = self._evaluators.get(key)
            if evaluator is None or evaluator.net is not pv_func:
                if self._cache_size > 0:    # device evaluation cache instead of the host LRU (same results, fewer network rows)
                    if self._dev_cache is None:
                        self._dev_cache = ds.EvalCache(self._game_name, self._cache_size, self.mcts._device)
                    evaluator = ds.CachedNetEvaluator(pv_func, self._dev_cache)
                else:
                    evaluator = ds.NetEvaluator(pv_func)
                self._evaluators = {key: evaluator}
            elif self._cache_size > 0:
                evaluator.net_rows = 0
        self._last_evaluator = evaluator
        if time_budget is None:
            ds.playout_device(self.mcts, buf, max_n, K, evaluator, stream)
            return
        # time-budgeted search (src/MCTS_cpp.py:112-128, 250-262: n_playout is the cap): the schedule is issued in chunks of a few
        # iterations with a stream synchronisation in between to read the clock (and, like the reference, stop early once the runner-up
        # can no longer catch up); the granularity of the budget is one chunk instead of one iteration
        sched = ds.iteration_schedule(max_n, K)
        t0, done, pos, chunk = time.perf_counter(), 0, 0, 4
        while pos < len(sched):
            part = sched[pos:pos + chunk]
            ds.playout_device(self.mcts, buf, 0, K, evaluator, stream, iters=part, shards=1)
            torch.cuda.current_stream().synchronize()
            pos += len(part)
            done += sum(max(k, 1) for k in part)
            elapsed = time.perf_counter() - t0
            if elapsed >= time_budget:
                break
            if self._should_early_exit(done, (time_budget - elapsed) / (elapsed / done)):
                break

    # ------------------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------------------
    def _predict_batch(self, pv_func, states, action_mask):
        return pv_func.predict(states, action_mask=action_mask)

    def _should_early_exit(self, step, remaining_steps):
        if step < 8:
            return False
        counts = np.array(self.mcts.get_all_counts()).reshape(self.batch_size, self.action_size)
        top2 = np.sort(counts, axis=1)[:, -2:]
        return bool(np.all(top2[:, 1] - top2[:, 0] > remaining_steps))

    def _evaluate_host(self, pv_func, leaf_boards, leaf_turns, is_term, term_d, term_p1w, term_p2w, valid_masks, use_cache):
        total = leaf_boards.shape[0]
        term_mask = is_term.astype(bool)
        d_vals, p1w_vals, p2w_vals = term_d.copy(), term_p1w.copy(), term_p2w.copy()
        moves_left = np.zeros(total, dtype=np.float32)
        probs = np.zeros((total, self.action_size), dtype=np.float32)
        idx = np.where(~term_mask)[0]
        if idx.size == 0:
            return probs, d_vals, p1w_vals, p2w_vals, moves_left
        miss = idx
        keys = None
        if use_cache and self.cache is not None:
            keys = {int(i): leaf_boards[i].tobytes() + int(leaf_turns[i]).to_bytes(1, "little", signed=True) for i in idx}
            miss = []
            for i in idx:
                k = keys[int(i)]
                if k in self.cache:
                    p, wdl, ml = self.cache.get(k)
                    probs[i] = p
                    d_vals[i] = wdl[0]
                    p1w_vals[i], p2w_vals[i] = (wdl[1], wdl[2]) if leaf_turns[i] == 1 else (wdl[2], wdl[1])
                    moves_left[i] = ml
                else:
                    miss.append(i)
            miss = np.asarray(miss, dtype=np.int64)
        if len(miss):
            conv = self._convert_board(leaf_boards[miss], leaf_turns[miss])
            masks = valid_masks[miss].astype(bool, copy=False)
            nn_probs, nn_wdl, nn_ml = self._predict_batch(pv_func, conv, masks)
            nn_ml = np.asarray(nn_ml).reshape(-1)
            probs[miss] = nn_probs
            d_abs, p1_abs, p2_abs = _relative_wdl_to_absolute(np.asarray(nn_wdl), leaf_turns[miss])
            d_vals[miss], p1w_vals[miss], p2w_vals[miss], moves_left[miss] = d_abs, p1_abs, p2_abs, nn_ml
            if keys is not None:
                for j, i in enumerate(miss):
                    k = keys[int(i)]
                    self.cache.put(k, (np.array(nn_probs[j]), np.array(nn_wdl[j]), float(nn_ml[j])))
                    self.cache._od[k]["state"] = conv[j:j + 1]
                    self.cache._od[k]["valid_mask"] = masks[j:j + 1].copy()
        f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
        return f32(probs), f32(d_vals), f32(p1w_vals), f32(p2w_vals), f32(moves_left)

    def batch_playout(self, pv_func, current_boards, turns, n_playout=None, vl_batch=1, time_budget=None):
        max_n = n_playout if n_playout is not None else self.n_playout
        use_time = time_budget is not None and time_budget > 0
        if hasattr(pv_func, "score_scale"):
            pv_func.score_scale = self.mcts.config.score_scale
        from . import device_search as ds
        if not self._custom_converter and ds.ReferenceNetAdapter.accepts(pv_func):
            # an unmodified network of the reference on a CUDA device: same numbers as its predict(), but nothing leaves the device
            ad = self._adapters.get(id(pv_func))
            if ad is None or ad.net is not pv_func:
                ad = ds.ReferenceNetAdapter(pv_func, self._game_name)
                self._adapters = {id(pv_func): ad}
            pv_func = ad
        if not self._custom_converter and (isinstance(pv_func, ds.SyntheticEvaluator) or hasattr(pv_func, "predict_device")):
            self._playout_device(pv_func, np.asarray(current_boards), np.asarray(turns), max_n, vl_batch, time_budget if use_time else None)
            return self
        current_boards = np.asarray(current_boards).astype(np.int8)
        turns = np.asarray(turns).astype(np.int32)
        t0 = time.perf_counter() if use_time else 0.0
        if vl_batch <= 1:
            for step in range(max_n):
                lb, td, tp1, tp2, it, lt, vm = self.mcts.search_batch(current_boards, turns)
                self.mcts.backprop_batch(*self._evaluate_host(pv_func, lb, lt, it, td, tp1, tp2, vm, True), it)
                if use_time:
                    done, elapsed = step + 1, time.perf_counter() - t0
                    if elapsed >= time_budget or self._should_early_exit(done, (time_budget - elapsed) / (elapsed / done)):
                        break
            return self
        K, remaining, total_sims = vl_batch, max_n, 0
        if remaining > 0:      # warm-up simulation bypasses the cache (src/MCTS_cpp.py:217-248)
            lb, td, tp1, tp2, it, lt, vm = self.mcts.search_batch(current_boards, turns)
            self.mcts.backprop_batch(*self._evaluate_host(pv_func, lb, lt, it, td, tp1, tp2, vm, False), it)
            remaining -= 1
            total_sims += 1
        while remaining > 0:
            if use_time:
                elapsed = time.perf_counter() - t0
                if elapsed >= time_budget:
                    break
                if total_sims > 0 and self._should_early_exit(total_sims, (time_budget - elapsed) / (elapsed / total_sims)):
                    break
            cur_K = min(K, remaining)
            remaining -= cur_K
            lb, td, tp1, tp2, it, lt, sym, vm = self.mcts.search_batch_vl(cur_K, current_boards, turns)
            try:
                ev = self._evaluate_host(pv_func, lb, lt, it, td, tp1, tp2, vm, True)
                self.mcts.backprop_batch_vl(cur_K, *ev, it, sym)
            except BaseException:
                self.mcts.remove_all_vl(cur_K)      # exception safety (src/MCTS_cpp.py:351-355)
                raise
            total_sims += cur_K
        return self

    def refresh_cache(self, pv_func):
        if self._dev_cache is not None:     # device cache: stale after a weight reload -> drop (misses re-evaluate lazily)
            self._dev_cache.clear()
        if self.cache is None or len(self.cache) == 0:
            return self
        if hasattr(pv_func, "score_scale"):
            pv_func.score_scale = self.mcts.config.score_scale
        od = self.cache._od
        keys = list(od.keys())
        states = np.concatenate([od[k]["state"] for k in keys], axis=0)
        masks = np.concatenate([od[k]["valid_mask"] for k in keys], axis=0) if all("valid_mask" in od[k] for k in keys) else None
        p, w, m = self._predict_batch(pv_func, states, masks)
        m = np.asarray(m).reshape(-1)
        for j, k in enumerate(keys):
            od[k]["value"] = (np.array(p[j]), np.array(w[j]), float(m[j]))
        return self

    def rollout_playout(self, current_boards, turns):
        if self._rollout_eval is None:
            self._rollout_eval = getattr(mcts_cpp, f"RolloutEvaluator_{self._game_name}")()
        self.mcts.search(self._rollout_eval, np.asarray(current_boards).astype(np.int8), np.asarray(turns).astype(np.int32),
                         self.n_playout)
        return self

    def set_noise_epsilon(self, eps):
        self.mcts.config.noise_epsilon = eps

    def set_mlh_params(self, slope, cap):
        self.mcts.config.mlh_slope, self.mcts.config.mlh_cap = slope, cap

    def set_score_utility_params(self, factor, scale):
        cfg = self.mcts.config
        old = cfg.score_scale
        cfg.score_utility_factor, cfg.score_scale = factor, scale
        if scale != old and self._dev_cache is not None:
            self._dev_cache.clear()
        if scale != old and self.cache is not None and len(self.cache) > 0:
            self.cache._od.clear()

    def set_c_init(self, val):
        self.mcts.config.c_init = val

    def set_c_base(self, val):
        self.mcts.config.c_base = val

    def set_alpha(self, val):
        self.mcts.config.dirichlet_alpha = val

    def set_fpu_reduction(self, val):
        self.mcts.config.fpu_reduction = val

    def set_use_symmetry(self, val):
        self.mcts.config.use_symmetry = val

    def set_value_decay(self, val):
        self.mcts.config.value_decay = val

    def reset_env(self, index):
        self.mcts.reset_env(index)
        return self

    def seed(self, seed):
        self.mcts.set_seed(seed)

    def prune_roots(self, actions):
        self.mcts.prune_roots(np.ascontiguousarray(actions, dtype=np.int32))
        return self

    def get_visits_count(self):
        return self.mcts.get_all_counts_array64()

    def get_mcts_probs(self):
        counts = self.get_visits_count()
        return counts / counts.sum(axis=1, keepdims=True)

    def get_root_stats(self):
        raw = self.mcts.get_all_root_stats()
        B, A = self.batch_size, self.action_size
        root, ch = raw[:, :6], raw[:, 6:].reshape(B, A, 8)
        out = {k: root[:, i] for i, k in enumerate(("root_N", "root_Q", "root_M", "root_D", "root_P1W", "root_P2W"))}
        out.update({k: ch[:, :, i] for i, k in enumerate(("N", "Q", "prior", "noise", "M", "D", "P1W", "P2W"))})
        return out
