"""Loader for libazb200.so (the CUDA kernels + C ABI declared in include/azb200.h).

There is NO CPU fallback: if the library is missing it is built with nvcc, and if that fails - or no CUDA
device is present when an engine is created - the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libazb200.so")
CSRC = os.path.join(_HERE, "csrc")


class AzSearchConfig(C.Structure):
    """az_search_config (include/azb200.h) == SearchConfig (src/cpp/MCTSNode.h:47-61)."""
    _fields_ = [(n, C.c_float) for n in (
        "c_init", "c_base", "dirichlet_alpha", "noise_epsilon", "fpu_reduction", "mlh_slope", "mlh_cap",
        "score_utility_factor", "score_scale", "value_decay")] + [("use_symmetry", C.c_int32), ("vl_count", C.c_int32)]


class AzHostLeaves(C.Structure):
    """az_host_leaves (include/azb200.h): leaf arrays of a host search inside a pinned block the caller owns."""
    _fields_ = [("block", C.c_int32), ("rows", C.c_int32), ("boards", C.c_void_p), ("term_d", C.c_void_p), ("term_p1w", C.c_void_p),
                ("term_p2w", C.c_void_p), ("is_term", C.c_void_p), ("turns", C.c_void_p), ("sym_ids", C.c_void_p), ("valid_mask", C.c_void_p)]


class AzRoot(C.Structure):
    """az_root (include/azb200.h): one position as bitboards, 32 bytes."""
    _fields_ = [("bb0", C.c_uint64), ("bb1", C.c_uint64), ("turn", C.c_int32), ("passes", C.c_int32), ("last", C.c_int32),
                ("reserved", C.c_int32)]


class AzGomoku(C.Structure):
    """az_gomoku (include/azb200_gomoku.h): one Gomoku position as row bit masks, 288 bytes."""
    _fields_ = [("rows", (C.c_uint32 * 32) * 2)] + [(n, C.c_int32) for n in (
        "size", "n_in_row", "turn", "n_pieces", "last_action", "last_player", "winner", "done")]


def build(force: bool = False) -> str:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-j8", "-C", CSRC] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libazb200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("AZB200_LIB") or LIB_PATH          # AZB200_LIB: another build of the same sources (A/B measurements)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.az_version.restype = C.c_char_p
    L.az_global_last_error.restype = C.c_char_p
    L.az_mcts_last_error.restype = C.c_char_p
    L.az_mcts_last_error.argtypes = [_vp]
    L.az_mcts_create.restype = _vp
    L.az_mcts_create.argtypes = [_i, _i, _i]
    L.az_mcts_destroy.restype = None
    L.az_mcts_destroy.argtypes = [_vp]
    L.az_search_config_defaults.restype = None
    L.az_mcts_compactions.restype = C.c_uint64
    L.az_mcts_compactions.argtypes = [_vp]
    sig = {
        "az_mcts_num_envs": [_vp],
        "az_mcts_set_config": [_vp, _vp], "az_mcts_get_config": [_vp, _vp],
        "az_mcts_set_seed": [_vp, _i64], "az_mcts_reset_env": [_vp, _i], "az_mcts_prune_roots": [_vp, _vp],
        "az_mcts_search_batch": [_vp] + [_vp] * 9,
        "az_mcts_backprop_batch": [_vp] + [_vp] * 6,
        "az_mcts_remove_all_vl": [_vp, _i], "az_mcts_search_batch_pinned": [_vp, _i, _vp, _vp, _vp], "az_pinned_release": [_i],
        "az_mcts_search_batch_vl": [_vp, _i] + [_vp] * 10,
        "az_mcts_backprop_batch_vl": [_vp, _i] + [_vp] * 7,
        "az_mcts_search": [_vp, _i, _vp, _vp, _i],
        "az_mcts_get_counts": [_vp, _vp], "az_mcts_get_counts64": [_vp, _vp], "az_mcts_get_root_stats": [_vp, _vp],
        "az_mcts_prune_roots_dev": [_vp, _vp, _vp], "az_mcts_reset_all_dev": [_vp, _vp],
        "az_mcts_search_dev": [_vp, _i, _vp, _vp, _vp],
        "az_pack_roots_dev": [_i, _i, _vp, _vp, _vp, _vp],
        "az_unpack_leaves_dev": [_i, _i] + [_vp] * 11,
        "az_mcts_set_env_base": [_vp, C.c_uint64], "az_selfplay_pos_bytes": [_i], "az_selfplay_max_plies": [_i],
        "az_selfplay_expand_dev": [_i, _i, _vp, _vp, _i] + [_vp] * 9,
        "az_selfplay_ply_dev": [_vp, _vp, _vp, _vp], "az_selfplay_flush_dev": [_vp, _vp],
        "az_mcts_set_lanes": [_vp, _i], "az_mcts_get_lanes": [_vp], "az_mcts_reserve": [_vp, _i],
        "az_mcts_set_lazy": [_vp, _i], "az_mcts_get_lazy": [_vp], "az_mcts_set_variant": [_vp, _i], "az_mcts_get_variant": [_vp], "az_mcts_set_wave_max": [_vp, _i], "az_mcts_get_wave_max": [_vp], "az_selftest_div": [_i, C.c_uint64, C.c_uint64, _vp],
        "az_mcts_backprop_dev": [_vp, _i] + [_vp] * 8,
        "az_mcts_search_range_dev": [_vp, _i, _vp, _vp, _i, _i, _i64, _i, _vp],
        "az_mcts_backprop_range_dev": [_vp, _i] + [_vp] * 7 + [_i, _i, _i64, _vp],
        "az_mcts_stream_handover_dev": [_vp, _vp],
        "az_mcts_playout_synthetic_dev": [_vp, _i, _i, _i, _i] + [_vp] * 9,
        "az_mcts_playout_synthetic_host": [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp], "az_mcts_get_counts64_pinned": [_vp, _vp, _vp],
        "az_mcts_search_eval_dev": [_vp, _i, _vp, _i, _vp],
        "az_mcts_get_counts_dev": [_vp, _vp, _vp], "az_mcts_get_root_stats_dev": [_vp, _vp, _vp],
        "az_mcts_enable_stats": [_vp, _i], "az_mcts_get_stats": [_vp, _vp], "az_mcts_get_warp_times": [_vp, _vp, _i], "az_mcts_time_select": [_vp, _i], "az_mcts_set_compaction": [_vp, _i], "az_mcts_get_select_time": [_vp, _vp, _vp, _vp], "az_mcts_get_backprop_time": [_vp, _vp, _vp, _vp],
        "az_eval_synthetic_dev": [_i, _i, _i] + [_vp] * 7,
        "az_eval_finalize_dev": [_i] + [_vp] * 8,
        "az_game_action_size": [_i], "az_game_board_size": [_i], "az_game_board_rows": [_i], "az_game_board_cols": [_i],
        "az_game_num_symmetries": [_i],
    }
    _u64 = C.c_uint64
    sig.update({
        "az_env_n_pieces": [_i, _vp], "az_env_winner": [_i, _vp], "az_env_full": [_i, _vp], "az_env_done": [_i, _vp],
        "az_env_valid_moves": [_i, _vp, _vp], "az_env_inverse_symmetry_action": [_i, _i, _i],
        "az_envs_reset_dev": [_i, _i, _vp, _vp], "az_envs_step_dev": [_i, _i, _vp, _vp, _vp, _vp, _vp],
        "az_envs_observe_dev": [_i, _i] + [_vp] * 7,
        "az_envs_rollout_dev": [_i, _i, _u64, _u64, _vp, _vp, _i, _i] + [_vp] * 7,
    })
    sig.update({
        "az_gomoku_set_params": [_vp, _i, _i], "az_gomoku_import": [_vp, _vp], "az_gomoku_step": [_vp, _i],
        "az_gomoku_valid_moves": [_vp, _vp], "az_gomoku_apply_symmetry": [_vp, _i],
        "az_gomoku_inverse_symmetry_action": [_i, _i, _i],
        "az_gomoku_reset_dev": [_i, _i, _i, _vp, _vp], "az_gomoku_step_dev": [_i] + [_vp] * 6,
        "az_gomoku_observe_dev": [_i, _i] + [_vp] * 7, "az_gomoku_symmetry_dev": [_i, _vp, _vp, _vp],
        "az_gomoku_rollout_dev": [_i, _i, _i, _u64, _u64, _vp, _vp, _i] + [_vp] * 7,
    })
    L.az_evalcache_create.restype = _vp
    L.az_evalcache_create.argtypes = [_i, _i, _i]
    L.az_evalcache_destroy.restype = None
    L.az_evalcache_destroy.argtypes = [_vp]
    sig.update({"az_evalcache_clear_dev": [_vp, _vp], "az_evalcache_lookup_dev": [_vp, _i] + [_vp] * 7,
                "az_evalcache_insert_dev": [_vp, _i] + [_vp] * 9, "az_evalcache_stats": [_vp, _vp],
                "az_evalcache_lookup_dedup_dev": [_vp, _i] + [_vp] * 8, "az_evalcache_resolve_dups_dev": [_vp, _i] + [_vp] * 5,
                "az_evalcache_dups": [_vp, _vp]})
    for name, argt in (("az_env_reset", [_i, _vp]), ("az_env_import", [_i, _vp, _vp]), ("az_env_export", [_i, _vp, _vp]),
                       ("az_env_step", [_i, _vp, _i]), ("az_env_apply_symmetry", [_i, _vp, _i]),
                       ("az_gomoku_reset", [_vp]), ("az_gomoku_export", [_vp, _vp])):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = argt
    for name, argt in sig.items():
        f = getattr(L, name, None)
        if f is None and os.environ.get("AZB200_LIB"):        # an older build in an A/B run may lack the newest entry points
            continue
        f = getattr(L, name)
        f.restype = _i
        f.argtypes = argt
    _lib = L
    return L
