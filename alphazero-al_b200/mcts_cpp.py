"""Drop-in replacement for the reference's pybind module ``src.mcts_cpp`` (src/cpp/mcts_bindings.cpp:372-395).

Same Python surface - ``SearchConfig``, ``BatchedMCTS_Connect4``, ``BatchedMCTS_Othello``, ``IEvaluator_<G>``,
``RolloutEvaluator_<G>`` with the same method names, argument names, dtypes, return tuples and RuntimeError
behaviour - so ``src/MCTS_cpp.py``, ``src/player.py``, ``src/pipeline.py`` and ``client.py`` run on it unmodified
(copy or symlink this file as ``src/mcts_cpp.py``).  Every call goes through the C ABI of
libazb200.so (include/azb200.h) into hand-written sm_100a kernels; trees never leave HBM and there is no CPU
fallback.

Extra, B200-only surface: the ``*_dev`` methods take CUDA device pointers (e.g. ``torch.Tensor.data_ptr()``)
and a stream, never synchronise, and can encode leaves straight into the CNN input tensor.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import AzHostLeaves, AzSearchConfig

GAME_IDS = {"Connect4": 0, "Othello": 1}
EVAL_UNIFORM, EVAL_ROLLOUT = 0, 1


class SearchConfig:
    """mcts_bindings.cpp:377-390 - plain value object with the 12 search parameters (MCTSNode.h:47-61)."""
    __slots__ = ("_c",)
    _FLOATS = ("c_init", "c_base", "dirichlet_alpha", "noise_epsilon", "fpu_reduction", "mlh_slope", "mlh_cap",
               "score_utility_factor", "score_scale", "value_decay")

    def __init__(self):
        object.__setattr__(self, "_c", AzSearchConfig())
        _lib.lib().az_search_config_defaults(C.byref(self._c))

    def __getattr__(self, name):
        if name in SearchConfig._FLOATS:
            return getattr(self._c, name)
        if name == "use_symmetry":
            return bool(self._c.use_symmetry)
        if name == "vl_count":
            return int(self._c.vl_count)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in SearchConfig._FLOATS:
            setattr(self._c, name, float(value))
        elif name == "use_symmetry":
            self._c.use_symmetry = 1 if value else 0
        elif name == "vl_count":
            self._c.vl_count = int(value)
        else:
            raise AttributeError(f"SearchConfig has no attribute {name!r}")

    def _copy_from(self, other: "SearchConfig"):
        C.memmove(C.byref(self._c), C.byref(other._c), C.sizeof(AzSearchConfig))

    def __repr__(self):
        f = ", ".join(f"{n}={getattr(self, n)!r}" for n in SearchConfig._FLOATS + ("use_symmetry", "vl_count"))
        return f"SearchConfig({f})"


class _IEvaluator:
    """IEvaluator_<Game> - abstract in Python (no constructor), mcts_bindings.cpp:46."""
    _kind = EVAL_UNIFORM

    def __init__(self, *a, **k):
        raise TypeError(f"{type(self).__name__}: No constructor defined!")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _carr(x, dtype):
    """py::array_t<T, c_style | forcecast>: silently cast + make contiguous (mcts_bindings.cpp:73,90-91)."""
    return np.ascontiguousarray(x, dtype=dtype)


class _BatchedMCTS:
    _game = None           # set by subclasses
    action_size = 0
    board_size = 0
    board_shape = ()

    def __init__(self, n_envs: int, device: int | None = None):
        L = _lib.lib()
        self._L = L
        self._n = int(n_envs)
        if device is None:
            device = _current_device()
        self._device = int(device)
        self._h = L.az_mcts_create(GAME_IDS[self._game], self._n, int(device))
        if not self._h:
            raise RuntimeError("BatchedMCTS_%s: %s" % (self._game, L.az_global_last_error().decode()))
        self._h = C.c_void_p(self._h)
        self._cfg = SearchConfig()
        self._A = self.action_size
        self._S = self.board_size

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._L.az_mcts_destroy(h)
            except Exception:
                pass

    # -- helpers ------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self._L.az_mcts_last_error(self._h).decode())

    def _push_cfg(self):
        self._ck(self._L.az_mcts_set_config(self._h, C.byref(self._cfg._c)))

    # -- config (reference_internal getter, copying setter: mcts_bindings.cpp:55-58) -----------
    @property
    def config(self) -> SearchConfig:
        return self._cfg

    @config.setter
    def config(self, cfg: SearchConfig):
        self._cfg._copy_from(cfg)

    def set_seed(self, seed: int):
        self._ck(self._L.az_mcts_set_seed(self._h, int(seed)))

    def reset_env(self, env_idx: int):
        self._ck(self._L.az_mcts_reset_env(self._h, int(env_idx)))

    def get_num_envs(self) -> int:
        return self._n

    def prune_roots(self, actions):
        a = _carr(actions, np.int32)
        if a.ndim != 1:
            raise RuntimeError("Actions must be 1D array")
        if a.size != self._n:
            raise RuntimeError(f"prune_roots: actions size ({a.size}) must match n_envs ({self._n})")
        self._push_cfg()
        self._ck(self._L.az_mcts_prune_roots(self._h, _ptr(a)))

    # -- AlphaZero split search (NN evaluation in Python) ---------------------------------------
    def _leaf_arrays(self, total):
        return (np.empty((total, *self.board_shape), np.int8), np.empty(total, np.float32), np.empty(total, np.float32),
                np.empty(total, np.float32), np.empty(total, np.uint8), np.empty(total, np.int32),
                np.empty((total, self._A), np.uint8))

    def _search_pinned(self, K, b, t):
        """One call, one device-to-host copy, no second pass: the returned arrays are views of a pinned block that goes back to the
        library's pool when the last of them is garbage collected (fresh arrays owned by Python, as mcts_bindings.cpp returns)."""
        out = AzHostLeaves()
        self._push_cfg()
        self._ck(self._L.az_mcts_search_batch_pinned(self._h, K, _ptr(b), _ptr(t), C.byref(out)))
        rows, A, S = out.rows, self._A, self._S
        base = out.boards                                         # lowest address of the block's arrays
        end = max(out.valid_mask + rows * A, out.sym_ids + rows * 4, out.turns + rows * 4, out.is_term + rows)
        buf = (C.c_uint8 * (end - base)).from_address(base)
        weakref.finalize(buf, self._L.az_pinned_release, out.block)
        raw = np.frombuffer(buf, dtype=np.uint8)

        def arr(ptr, dtype, count, shape):
            o = ptr - base
            return raw[o:o + count * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
        ob = arr(out.boards, np.int8, rows * S, (rows, *self.board_shape))
        td, tp1, tp2 = (arr(p, np.float32, rows, (rows,)) for p in (out.term_d, out.term_p1w, out.term_p2w))
        it = arr(out.is_term, np.uint8, rows, (rows,))
        ot = arr(out.turns, np.int32, rows, (rows,))
        sym = arr(out.sym_ids, np.int32, rows, (rows,))
        vm = arr(out.valid_mask, np.uint8, rows * A, (rows, A))
        return ob, td, tp1, tp2, it, ot, sym, vm

    def search_batch(self, input_boards, turns):
        b, t = _carr(input_boards, np.int8), _carr(turns, np.int32)
        batch = b.shape[0] if b.ndim else 0
        if batch != self._n:
            raise RuntimeError(f"search_batch: input_boards batch size ({batch}) must match n_envs ({self._n})")
        if t.size != batch:
            raise RuntimeError("Turns size must match batch size")
        if b.size != batch * self._S:
            raise RuntimeError(f"search_batch: input_boards must have {self._S} cells per board")
        ob, td, tp1, tp2, it, ot, _, vm = self._search_pinned(0, b, t)
        return ob, td, tp1, tp2, it, ot, vm

    def backprop_batch(self, policy_logits, d_vals, p1w_vals, p2w_vals, moves_left, is_term):
        pol = _carr(policy_logits, np.float32)
        d, p1, p2, ml = (_carr(x, np.float32) for x in (d_vals, p1w_vals, p2w_vals, moves_left))
        it = _carr(is_term, np.uint8)
        n = self._n
        rows = pol.shape[0] if pol.ndim else 0
        if rows != n:
            raise RuntimeError(f"backprop_batch: policy_logits batch size ({rows}) must match n_envs ({n})")
        if d.size != n or p1.size != n or p2.size != n:
            raise RuntimeError(f"backprop_batch: d/p1w/p2w size must match n_envs ({n})")
        if ml.size != n:
            raise RuntimeError(f"backprop_batch: moves_left size ({ml.size}) must match n_envs ({n})")
        if it.size != n:
            raise RuntimeError(f"backprop_batch: is_term size ({it.size}) must match n_envs ({n})")
        if pol.size != n * self._A:
            raise RuntimeError(f"backprop_batch: policy_logits must have {self._A} columns")
        self._push_cfg()
        self._ck(self._L.az_mcts_backprop_batch(self._h, _ptr(pol), _ptr(d), _ptr(p1), _ptr(p2), _ptr(ml), _ptr(it)))

    def remove_all_vl(self, K):
        self._push_cfg()
        self._ck(self._L.az_mcts_remove_all_vl(self._h, int(K)))

    def search_batch_vl(self, K, input_boards, turns):
        b, t = _carr(input_boards, np.int8), _carr(turns, np.int32)
        n = self._n
        batch = b.shape[0] if b.ndim else 0
        if batch != n:
            raise RuntimeError(f"search_batch_vl: input batch ({batch}) != n_envs ({n})")
        if t.size != n:
            raise RuntimeError("search_batch_vl: turns size must match n_envs")
        K = int(K)
        if K < 1:
            raise RuntimeError("search_batch_vl: K must be >= 1")
        if b.size != batch * self._S:
            raise RuntimeError(f"search_batch_vl: input_boards must have {self._S} cells per board")
        return self._search_pinned(K, b, t)

    def backprop_batch_vl(self, K, policy_logits, d_vals, p1w_vals, p2w_vals, moves_left, is_term, sym_ids):
        K = int(K)
        pol = _carr(policy_logits, np.float32)
        d, p1, p2, ml = (_carr(x, np.float32) for x in (d_vals, p1w_vals, p2w_vals, moves_left))
        it, sym = _carr(is_term, np.uint8), _carr(sym_ids, np.int32)
        total = self._n * K
        rows = pol.shape[0] if pol.ndim else 0
        if rows != total:
            raise RuntimeError(f"backprop_batch_vl: policy batch ({rows}) != N*K ({total})")
        if d.size != total or p1.size != total or p2.size != total:
            raise RuntimeError("backprop_batch_vl: d/p1w/p2w size must be N*K")
        if ml.size != total:
            raise RuntimeError("backprop_batch_vl: moves_left size must be N*K")
        if it.size != total:
            raise RuntimeError("backprop_batch_vl: is_term size must be N*K")
        if sym.size != total:
            raise RuntimeError("backprop_batch_vl: sym_ids size must be N*K")
        if pol.size != total * self._A:
            raise RuntimeError(f"backprop_batch_vl: policy_logits must have {self._A} columns")
        self._push_cfg()
        self._ck(self._L.az_mcts_backprop_batch_vl(self._h, K, _ptr(pol), _ptr(d), _ptr(p1), _ptr(p2), _ptr(ml),
                                                   _ptr(it), _ptr(sym)))

    # -- whole loop on device with a built-in evaluator (mcts_bindings.cpp:313-337) --------------
    def search(self, evaluator, input_boards, turns, n_playout):
        if not isinstance(evaluator, self._ievaluator):
            raise TypeError("search(): incompatible function arguments (evaluator must be an IEvaluator_%s)" % self._game)
        b, t = _carr(input_boards, np.int8), _carr(turns, np.int32)
        batch = b.shape[0] if b.ndim else 0
        if batch != self._n:
            raise RuntimeError(f"search: input_boards batch size ({batch}) must match n_envs ({self._n})")
        if t.size != batch:
            raise RuntimeError("search: turns size must match batch size")
        self._push_cfg()
        self._ck(self._L.az_mcts_search(self._h, evaluator._kind, _ptr(b), _ptr(t), int(n_playout)))

    # -- statistics ---------------------------------------------------------------------------------
    def get_all_counts(self):
        out = np.empty(self._n * self._A, np.int32)
        self._ck(self._L.az_mcts_get_counts(self._h, _ptr(out)))
        return out.tolist()          # std::vector<int> -> list (mcts_bindings.cpp:342)

    def get_all_counts_array(self):
        """B200-only convenience: the same counts as an int32 ndarray [n_envs, A] (no Python list)."""
        out = np.empty((self._n, self._A), np.int32)
        self._ck(self._L.az_mcts_get_counts(self._h, _ptr(out)))
        return out

    def get_all_counts_array64(self):
        """The counts as int64 [n_envs, A] - the dtype np.array(get_all_counts()) has in the reference wrapper.  The array is a view
        of a pinned block (the destination of the one device-to-host copy, or the block playout_synthetic_host already filled) that
        goes back to the library's pool when the array is garbage collected."""
        ptr, block = C.c_void_p(), C.c_int(-1)
        self._ck(self._L.az_mcts_get_counts64_pinned(self._h, C.byref(ptr), C.byref(block)))
        buf = (C.c_int64 * (self._n * self._A)).from_address(ptr.value)
        weakref.finalize(buf, self._L.az_pinned_release, block.value)
        return np.frombuffer(buf, dtype=np.int64).reshape(self._n, self._A)

    def get_all_root_stats(self):
        out = np.empty((self._n, 6 + 8 * self._A), np.float32)
        self._ck(self._L.az_mcts_get_root_stats(self._h, _ptr(out)))
        return out

    # -- B200-only: device-pointer twins (no copies, no synchronisation) ---------------------------
    # roots_ptr -> az_root[n], leaves_ptr -> az_leaf[n*max(K,1)] (32-byte records, include/azb200.h)
    def search_dev(self, K, roots_ptr, leaves_ptr, stream=0):
        self._push_cfg()
        self._ck(self._L.az_mcts_search_dev(self._h, int(K), roots_ptr, leaves_ptr, stream or None))

    def backprop_dev(self, K, policy, d, p1w, p2w, ml, is_term=0, sym=0, stream=0):
        self._push_cfg()
        self._ck(self._L.az_mcts_backprop_dev(self._h, int(K), policy, d, p1w, p2w, ml, is_term or None, sym or None,
                                              stream or None))

    def search_range_dev(self, K, roots_ptr, leaves_ptr, first, count, row0, new_epoch=True, stream=0):
        """search_dev restricted to trees [first, first + count); the shard's rows start at row0 (include/azb200.h)."""
        self._push_cfg()
        self._ck(self._L.az_mcts_search_range_dev(self._h, int(K), roots_ptr, leaves_ptr, int(first), int(count), int(row0),
                                                  1 if new_epoch else 0, stream or None))

    def backprop_range_dev(self, K, policy, d, p1w, p2w, ml, first, count, row0, is_term=0, sym=0, stream=0):
        self._push_cfg()
        self._ck(self._L.az_mcts_backprop_range_dev(self._h, int(K), policy, d, p1w, p2w, ml, is_term or None, sym or None,
                                                    int(first), int(count), int(row0), stream or None))

    def playout_synthetic_dev(self, mode, n_playout, K, shards, roots_ptr, leaves_ptr, policy, d, p1w, p2w, ml, stream=0):
        """The whole playout loop with a synthetic evaluator, driven natively (include/azb200.h).  Returns kernel launches."""
        self._push_cfg()
        out = C.c_int(0)
        self._ck(self._L.az_mcts_playout_synthetic_dev(self._h, int(mode), int(n_playout), int(K), int(shards), roots_ptr, leaves_ptr,
                                                       policy, d, p1w, p2w, ml, stream or None, C.byref(out)))
        return out.value

    def playout_synthetic_host(self, mode, n_playout, K, shards, input_boards, turns, want_counts=True):
        """The same loop from host arrays, pipelined shard by shard (staging, copies, pack, loop and visit counts on the shard's own
        stream; include/azb200.h).  Synchronous.  With want_counts the next get_all_counts_array64() is free.  Returns kernel launches."""
        b = _carr(input_boards, np.int8)
        t = _carr(turns, np.int32)
        if b.size != self._n * self._S or t.size != self._n:
            raise RuntimeError(f"playout: boards / turns must hold n_envs ({self._n}) games")
        self._push_cfg()
        out = C.c_int(0)
        self._ck(self._L.az_mcts_playout_synthetic_host(self._h, int(mode), int(n_playout), int(K), int(shards), _ptr(b), _ptr(t),
                                                        1 if want_counts else 0, C.byref(out)))
        return out.value

    def set_compaction(self, mode):
        """Arena compaction at re-roots: 0 never, 1 auto (default), 2 always (include/azb200.h)."""
        self._ck(self._L.az_mcts_set_compaction(self._h, int(mode)))

    def compactions(self):
        return int(self._L.az_mcts_compactions(self._h))

    def time_select(self, on=True):
        """CUDA events around every select launch (measurement, include/azb200.h)."""
        self._ck(self._L.az_mcts_time_select(self._h, 1 if on else 0))

    def get_select_time(self):
        """(summed select ms, launches, leaf rows) since the last call; synchronises the device."""
        ms, n, rows = C.c_float(0), C.c_int(0), C.c_uint64(0)
        self._ck(self._L.az_mcts_get_select_time(self._h, C.byref(ms), C.byref(n), C.byref(rows)))
        return ms.value, n.value, rows.value

    def get_backprop_time(self):
        """(summed back-prop ms, launches, leaf rows) since the last call; synchronises the device."""
        ms, n, rows = C.c_float(0), C.c_int(0), C.c_uint64(0)
        self._ck(self._L.az_mcts_get_backprop_time(self._h, C.byref(ms), C.byref(n), C.byref(rows)))
        return ms.value, n.value, rows.value

    def stream_handover_dev(self, stream=0):
        self._ck(self._L.az_mcts_stream_handover_dev(self._h, stream or None))

    def prune_roots_dev(self, actions_ptr, stream=0):
        self._push_cfg()
        self._ck(self._L.az_mcts_prune_roots_dev(self._h, actions_ptr, stream or None))

    def reset_all_dev(self, stream=0):
        """Every tree back to a fresh root, stream-ordered, host arena bookkeeping included (include/azb200.h)."""
        self._ck(self._L.az_mcts_reset_all_dev(self._h, stream or None))

    def search_eval_dev(self, evaluator_kind, roots_ptr, n_playout, stream=0):
        self._push_cfg()
        self._ck(self._L.az_mcts_search_eval_dev(self._h, int(evaluator_kind), roots_ptr, int(n_playout), stream or None))

    def get_counts_dev(self, out_ptr, stream=0):
        self._ck(self._L.az_mcts_get_counts_dev(self._h, out_ptr, stream or None))

    def get_root_stats_dev(self, out_ptr, stream=0):
        self._ck(self._L.az_mcts_get_root_stats_dev(self._h, out_ptr, stream or None))

    def set_lanes(self, lanes):
        """Lanes cooperating on one tree (Connect4: 1/2/4/8, 0 = auto from n_envs)."""
        self._ck(self._L.az_mcts_set_lanes(self._h, int(lanes)))

    def get_lanes(self):
        return self._L.az_mcts_get_lanes(self._h)

    def set_lazy(self, on=True):
        """Lazy edge blocks (include/azb200.h): header-only expansions, materialised on the second visit.  Off by default."""
        self._ck(self._L.az_mcts_set_lazy(self._h, 1 if on else 0))

    def get_lazy(self):
        return bool(self._L.az_mcts_get_lazy(self._h))

    def set_variant(self, variant):
        """Generation of the thread-per-tree Connect4 kernels (0 first, 1 lean); bit-identical results."""
        self._ck(self._L.az_mcts_set_variant(self._h, int(variant)))

    def get_variant(self):
        return self._L.az_mcts_get_variant(self._h)

    def set_wave_max(self, max_lanes):
        """Batches of at most `max_lanes` descent lanes (trees x 4 for K <= 4, x 8 for K <= 8) use the staggered-descent select
        (one lane per virtual-loss descent); 0 = off, default 131072."""
        self._ck(self._L.az_mcts_set_wave_max(self._h, int(max_lanes)))

    def get_wave_max(self):
        return self._L.az_mcts_get_wave_max(self._h)

    def set_env_base(self, base):
        """Global index of env 0 (keeps RNG streams sharding-invariant across GPUs)."""
        self._ck(self._L.az_mcts_set_env_base(self._h, int(base)))

    def reserve(self, slots_per_tree):
        self._ck(self._L.az_mcts_reserve(self._h, int(slots_per_tree)))

    def enable_stats(self, on=True):
        self._ck(self._L.az_mcts_enable_stats(self._h, 1 if on else 0))

    def get_stats(self):
        out = np.zeros(8, np.uint64)
        self._ck(self._L.az_mcts_get_stats(self._h, _ptr(out)))
        keys = ("sims", "depth", "edges_scanned", "edges_created", "expansions", "max_arena_slots", "arena_cap", "launches")
        return dict(zip(keys, (int(v) for v in out)))


def _current_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _make(game: str):
    L = _lib.lib()
    gid = GAME_IDS[game]
    ie = type(f"IEvaluator_{game}", (_IEvaluator,), {"_kind": EVAL_UNIFORM})

    def _re_init(self):
        pass
    re = type(f"RolloutEvaluator_{game}", (ie,), {"_kind": EVAL_ROLLOUT, "__init__": _re_init})
    cls = type(f"BatchedMCTS_{game}", (_BatchedMCTS,), {
        "_game": game,
        "_ievaluator": ie,
        "action_size": L.az_game_action_size(gid),
        "board_size": L.az_game_board_size(gid),
        "board_shape": (L.az_game_board_rows(gid), L.az_game_board_cols(gid)),
    })
    return cls, ie, re


BatchedMCTS_Connect4, IEvaluator_Connect4, RolloutEvaluator_Connect4 = _make("Connect4")
BatchedMCTS_Othello, IEvaluator_Othello, RolloutEvaluator_Othello = _make("Othello")

__all__ = ["SearchConfig", "BatchedMCTS_Connect4", "BatchedMCTS_Othello", "IEvaluator_Connect4",
           "IEvaluator_Othello", "RolloutEvaluator_Connect4", "RolloutEvaluator_Othello"]
