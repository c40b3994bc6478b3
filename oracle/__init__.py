"""Parity checker for the CUDA engine - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``alphazero-al_b200``) never does.

Two checkers live here:

* ``OracleMCTS`` / ``OracleEnv`` - ctypes front-end to ``az_oracle.c`` (our plain-C restatement of the
  reference algorithm, every function citing the reference file:line).
* ``load_ref(kind)`` - imports the UNMODIFIED reference engine compiled by ``oracle/Makefile`` into
  ``oracle/_ref/<kind>/`` (``parity``: -O2 -ffp-contract=off, ``timing``: -O3 -march=x86-64-v3).  The
  reference sources are compiled where they lie under /root/reference; only the built ``.so`` files exist
  here and they are git-ignored.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GAMES = {"Connect4": 0, "Othello": 1}
ACTION_SIZE = {"Connect4": 7, "Othello": 65}
BOARD_SHAPE = {"Connect4": (6, 7), "Othello": (8, 8)}
EVAL_UNIFORM, EVAL_ROLLOUT = 0, 1


class OrcConfig(C.Structure):
    """Mirror of ``orc_config`` (SearchConfig, src/cpp/MCTSNode.h:47-61)."""
    _fields_ = [(n, C.c_float) for n in (
        "c_init", "c_base", "dirichlet_alpha", "noise_epsilon", "fpu_reduction", "mlh_slope", "mlh_cap",
        "score_utility_factor", "score_scale", "value_decay")] + [("use_symmetry", C.c_int32), ("vl_count", C.c_int32)]


def build(force: bool = False) -> None:
    """Compile the C restatement and (when /root/reference is present) the reference build."""
    subprocess.run(["make", "-C", _HERE, "restate"] + (["-B"] if force else []), check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libaz_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE, "restate"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_int]
        _lib.orc_rollout_hash.restype = C.c_uint64
        _lib.orc_rollout_hash.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        for name in ("orc_destroy", "orc_set_config", "orc_get_config", "orc_set_seed", "orc_reset_env",
                     "orc_prune_roots", "orc_search_batch", "orc_backprop_batch", "orc_remove_all_vl",
                     "orc_search_batch_vl", "orc_backprop_batch_vl", "orc_search", "orc_get_counts",
                     "orc_get_root_stats", "orc_get_tree_stats", "orc_gmk_reset", "orc_gmk_symmetry", "orc_gmk_board",
                     "orc_gmk_fields", "orc_gmk_rollout_digests"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleMCTS:
    """Same method names / argument order / return tuples as the reference's ``mcts_cpp.BatchedMCTS_<Game>``
    (src/cpp/mcts_bindings.cpp:50-369), backed by the C restatement."""

    def __init__(self, game: str, n_envs: int):
        self.game = game
        self.gid = GAMES[game]
        self.n = n_envs
        self.A = ACTION_SIZE[game]
        self.shape = BOARD_SHAPE[game]
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(self.gid, n_envs))
        self.config = OrcConfig()
        self.L.orc_get_config(self.h, C.byref(self.config))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def _sync(self):
        self.L.orc_set_config(self.h, C.byref(self.config))

    def set_seed(self, seed: int):
        self.L.orc_set_seed(self.h, C.c_int64(seed))

    def reset_env(self, i: int):
        self.L.orc_reset_env(self.h, C.c_int(i))

    def get_num_envs(self):
        return self.n

    def prune_roots(self, actions):
        self._sync()
        a = np.ascontiguousarray(actions, dtype=np.int32)
        assert a.shape == (self.n,)
        self.L.orc_prune_roots(self.h, _p(a))

    def _outs(self, total):
        return (np.empty((total, *self.shape), np.int8), np.empty(total, np.float32), np.empty(total, np.float32),
                np.empty(total, np.float32), np.empty(total, np.uint8), np.empty(total, np.int32),
                np.empty((total, self.A), np.uint8))

    def search_batch(self, boards, turns):
        self._sync()
        b = np.ascontiguousarray(boards, dtype=np.int8)
        t = np.ascontiguousarray(turns, dtype=np.int32)
        ob, td, tp1, tp2, it, ot, vm = self._outs(self.n)
        self.L.orc_search_batch(self.h, _p(b), _p(t), _p(ob), _p(td), _p(tp1), _p(tp2), _p(it), _p(ot), _p(vm))
        return ob, td, tp1, tp2, it, ot, vm

    def backprop_batch(self, policy_logits, d_vals, p1w_vals, p2w_vals, moves_left, is_term):
        self._sync()
        f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
        pol, d, p1, p2, ml = map(f, (policy_logits, d_vals, p1w_vals, p2w_vals, moves_left))
        it = np.ascontiguousarray(is_term, dtype=np.uint8)
        self.L.orc_backprop_batch(self.h, _p(pol), _p(d), _p(p1), _p(p2), _p(ml), _p(it))

    def remove_all_vl(self, K):
        self._sync()
        self.L.orc_remove_all_vl(self.h, C.c_int(K))

    def search_batch_vl(self, K, boards, turns):
        self._sync()
        b = np.ascontiguousarray(boards, dtype=np.int8)
        t = np.ascontiguousarray(turns, dtype=np.int32)
        ob, td, tp1, tp2, it, ot, vm = self._outs(self.n * K)
        sym = np.empty(self.n * K, np.int32)
        self.L.orc_search_batch_vl(self.h, C.c_int(K), _p(b), _p(t), _p(ob), _p(td), _p(tp1), _p(tp2), _p(it),
                                   _p(ot), _p(sym), _p(vm))
        return ob, td, tp1, tp2, it, ot, sym, vm

    def backprop_batch_vl(self, K, policy_logits, d_vals, p1w_vals, p2w_vals, moves_left, is_term, sym_ids):
        self._sync()
        f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
        pol, d, p1, p2, ml = map(f, (policy_logits, d_vals, p1w_vals, p2w_vals, moves_left))
        it = np.ascontiguousarray(is_term, dtype=np.uint8)
        sym = np.ascontiguousarray(sym_ids, dtype=np.int32)
        self.L.orc_backprop_batch_vl(self.h, C.c_int(K), _p(pol), _p(d), _p(p1), _p(p2), _p(ml), _p(it), _p(sym))

    def search(self, evaluator_kind, boards, turns, n_playout):
        self._sync()
        b = np.ascontiguousarray(boards, dtype=np.int8)
        t = np.ascontiguousarray(turns, dtype=np.int32)
        self.L.orc_search(self.h, C.c_int(evaluator_kind), _p(b), _p(t), C.c_int(n_playout))

    def get_all_counts(self):
        out = np.empty(self.n * self.A, np.int32)
        self.L.orc_get_counts(self.h, _p(out))
        return out.tolist()

    def get_all_root_stats(self):
        out = np.empty((self.n, 6 + 8 * self.A), np.float32)
        self.L.orc_get_root_stats(self.h, _p(out))
        return out

    def tree_stats(self):
        out = np.zeros(36, np.uint64)
        self.L.orc_get_tree_stats(self.h, _p(out))
        d = dict(zip(("sims", "depth", "edges_scanned", "edges_created", "nodes", "edges", "scanned_allocated", "scanned_visited",
                      "expansions"), out[:9].tolist()))
        d.update(level_nodes=out[9:17].tolist(), level_edges=out[17:25].tolist(), level_allocated=out[25:33].tolist(),
                 expanded_nodes=int(out[33]), expanded_revisited=int(out[34]), expanded_revisited_edges=int(out[35]))
        return d


class OracleEnv:
    """Single-game env over the C restatement (checker for env_cpp.<game>.Env)."""

    def __init__(self, game: str):
        self.game, self.gid = game, GAMES[game]
        self.A, self.shape = ACTION_SIZE[game], BOARD_SHAPE[game]
        self.L = lib()
        self.state = C.create_string_buffer(self.L.orc_env_sizeof())
        self.reset()

    def reset(self):
        self.L.orc_env_reset(self.gid, self.state)

    def import_board(self, board, turn):
        b = np.ascontiguousarray(board, dtype=np.int8)
        self.L.orc_env_import(self.gid, self.state, _p(b), C.c_int(int(turn)))

    def step(self, a):
        self.L.orc_env_step(self.gid, self.state, C.c_int(int(a)))

    def winner(self):
        return self.L.orc_env_winner(self.gid, self.state)

    def full(self):
        return bool(self.L.orc_env_full(self.gid, self.state))

    def done(self):
        return bool(self.L.orc_env_done(self.gid, self.state))

    def valid_moves(self):
        m = np.empty(self.A, np.int32)
        n = self.L.orc_env_valid(self.gid, self.state, _p(m))
        return m[:n].tolist()

    def apply_symmetry(self, s):
        self.L.orc_env_symmetry(self.gid, self.state, C.c_int(int(s)))

    @property
    def board(self):
        out = np.empty(self.shape, np.int8)
        self.L.orc_env_board(self.gid, self.state, _p(out))
        return out

    @property
    def turn(self):
        return self.L.orc_env_turn(self.state)


def env_rollout(game: str, seed: int, gidx: int, record: bool = True):
    """Config-2 lockstep random rollout of one game on the C restatement."""
    L = lib()
    gid, A, shape = GAMES[game], ACTION_SIZE[game], BOARD_SHAPE[game]
    maxp = 42 if game == "Connect4" else 128
    digest = C.c_uint64(0)
    if record:
        boards = np.zeros((maxp, *shape), np.int8)
        masks = np.zeros((maxp, A), np.uint8)
        turns = np.zeros(maxp, np.int32)
        actions = np.zeros(maxp, np.int32)
        winners = np.zeros(maxp, np.int32)
        dones = np.zeros(maxp, np.uint8)
        n = L.orc_env_rollout(gid, C.c_uint64(seed), C.c_uint64(gidx), maxp, _p(boards), _p(masks), _p(turns),
                              _p(actions), _p(winners), _p(dones), C.byref(digest))
        return dict(plies=n, boards=boards[:n], masks=masks[:n], turns=turns[:n], actions=actions[:n],
                    winners=winners[:n], dones=dones[:n], digest=digest.value)
    n = L.orc_env_rollout(gid, C.c_uint64(seed), C.c_uint64(gidx), maxp, None, None, None, None, None, None,
                          C.byref(digest))
    return dict(plies=n, digest=digest.value)


def env_rollout_digests(game: str, seed: int, first: int, n: int):
    """(digests uint64[n], plies int32[n]) of config-2 games [first, first+n) on the C restatement."""
    L = lib()
    dig = np.empty(n, np.uint64)
    pl = np.empty(n, np.int32)
    L.orc_env_rollout_digests(GAMES[game], C.c_uint64(seed), C.c_uint64(first), C.c_int(n), _p(dig), _p(pl))
    return dig, pl


class OracleGomoku:
    """Single Gomoku game over the byte-board C restatement (checker for env_cpp.gomoku.Env and BatchedGomoku)."""

    def __init__(self, size: int = 15, n_in_row: int = 5):
        self.L = lib()
        self.state = C.create_string_buffer(self.L.orc_gmk_sizeof())
        if self.L.orc_gmk_set_params(self.state, C.c_int(size), C.c_int(n_in_row)) != 0:
            raise RuntimeError("invalid Gomoku parameters")
        self.size, self.k = size, n_in_row

    def reset(self):
        self.L.orc_gmk_reset(self.state)

    def import_board(self, board):
        b = np.ascontiguousarray(board, dtype=np.int8)
        return self.L.orc_gmk_import(self.state, _p(b))

    def step(self, a):
        """0, or 1 / 2 / 3 = finished / out of range / occupied (the reference's three exceptions)."""
        return self.L.orc_gmk_step(self.state, C.c_int(int(a)))

    def valid_moves(self):
        m = np.empty(self.size * self.size, np.int32)
        n = self.L.orc_gmk_valid(self.state, _p(m))
        return m[:n].tolist()

    def apply_symmetry(self, s):
        self.L.orc_gmk_symmetry(self.state, C.c_int(int(s)))

    @property
    def board(self):
        out = np.empty((self.size, self.size), np.int8)
        self.L.orc_gmk_board(self.state, _p(out))
        return out

    def _fields(self):
        f = np.empty(6, np.int32)
        self.L.orc_gmk_fields(self.state, _p(f))
        return f

    turn = property(lambda self: int(self._fields()[0]))
    n_pieces = property(lambda self: int(self._fields()[1]))
    last_action = property(lambda self: int(self._fields()[2]))

    def winner(self):
        return int(self._fields()[4])

    def done(self):
        return bool(self._fields()[5])


def gomoku_rollout(size: int, n_in_row: int, seed: int, gidx: int, record: bool = True):
    """Lockstep random rollout of one Gomoku game on the C restatement (twin of az_gomoku_rollout_dev)."""
    L = lib()
    S = size * size
    digest = C.c_uint64(0)
    final = np.zeros((size, size), np.int8)
    if record:
        boards = np.zeros((S, size, size), np.int8)
        turns, actions, winners = (np.zeros(S, np.int32) for _ in range(3))
        dones = np.zeros(S, np.uint8)
        n = L.orc_gmk_rollout(size, n_in_row, C.c_uint64(seed), C.c_uint64(gidx), _p(boards), _p(turns), _p(actions),
                              _p(winners), _p(dones), C.byref(digest), _p(final))
        return dict(plies=n, boards=boards[:n], turns=turns[:n], actions=actions[:n], winners=winners[:n], dones=dones[:n],
                    digest=digest.value, final=final)
    n = L.orc_gmk_rollout(size, n_in_row, C.c_uint64(seed), C.c_uint64(gidx), None, None, None, None, None,
                          C.byref(digest), _p(final))
    return dict(plies=n, digest=digest.value, final=final)


def gomoku_rollout_digests(size: int, n_in_row: int, seed: int, first: int, n: int):
    """(digests uint64[n], plies int32[n]) of Gomoku rollout games [first, first+n) on the C restatement."""
    L = lib()
    dig = np.empty(n, np.uint64)
    pl = np.empty(n, np.int32)
    L.orc_gmk_rollout_digests(size, n_in_row, C.c_uint64(seed), C.c_uint64(first), C.c_int(n), _p(dig), _p(pl))
    return dig, pl


def ref_available(kind: str = "parity") -> bool:
    d = os.path.join(_HERE, "_ref", kind)
    return os.path.isdir(d) and any(f.startswith("mcts_cpp") for f in os.listdir(d))


_ref_cache = {}


def load_ref(kind: str = "parity"):
    """Import (mcts_cpp, env_cpp) of the unmodified reference build.  Only one kind per process
    (pybind11 registers the C++ types globally)."""
    if _ref_cache:
        if kind not in _ref_cache:
            raise RuntimeError("a different oracle/_ref build is already loaded in this process")
        return _ref_cache[kind]
    d = os.path.join(_HERE, "_ref", kind)
    mods = []
    for name in ("mcts_cpp", "env_cpp"):
        cand = [f for f in os.listdir(d) if f.startswith(name) and f.endswith(".so")]
        if not cand:
            raise ImportError(f"oracle/_ref/{kind}/{name}*.so missing - run `make -C oracle ref`")
        spec = importlib.util.spec_from_file_location(name, os.path.join(d, cand[0]))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules.setdefault(name, mod)            # lets pickle find the reference's Env classes
        for sub in ("connect4", "othello", "gomoku"):
            if hasattr(mod, sub):
                sys.modules.setdefault(f"{name}.{sub}", getattr(mod, sub))
        mods.append(mod)
    _ref_cache[kind] = tuple(mods)
    return _ref_cache[kind]
