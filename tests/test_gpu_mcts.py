"""GPU parity tests: the CUDA engine (through the C ABI, host-buffer entry points) against the C restatement
(oracle/az_oracle.c) and - when the compiled reference travelled with the snapshot - against the unmodified
reference engine (oracle/_ref/parity).  Bit-exact: leaf boards, terminal flags/values, turns, legal masks,
symmetry ids (vs the restatement, which shares the counter-based RNG), visit counts, root statistics."""
import importlib

import numpy as np
import pytest

import oracle
from harness import SERVER_DEFAULTS, compare_engines, counts, playout, random_positions, set_config

pytestmark = pytest.mark.gpu

OTH_CFG = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False,
               score_utility_factor=0.15, score_scale=8.0)


def _cuda(game, n):
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    return getattr(m, f"BatchedMCTS_{game}")(n)


def _orc(game, n):
    return oracle.OracleMCTS(game, n)


def _ref(game, n):
    if not oracle.ref_available("parity"):
        pytest.skip("oracle/_ref/parity not present")
    mcts_cpp, _ = oracle.load_ref("parity")
    return getattr(mcts_cpp, f"BatchedMCTS_{game}")(n)


@pytest.mark.parametrize("K", [1, 2, 4, 8])
def test_c4_fresh_roots_vs_restatement(K):
    c = compare_engines(_cuda("Connect4", 64), _orc("Connect4", 64), "Connect4", 64, 60, K, SERVER_DEFAULTS)
    assert (c.sum(axis=1) == 59).all()


def test_c4_tree_reuse_decay_remainder_vs_restatement():
    cfg = dict(SERVER_DEFAULTS, value_decay=0.97)
    boards, turns = random_positions("Connect4", 96, 20, 1)
    compare_engines(_cuda("Connect4", 96), _orc("Connect4", 96), "Connect4", 96, 51, 4, cfg, boards=boards, turns=turns, moves=14)


def test_c4_symmetry_on_same_sym_stream_as_restatement():
    # the restatement and the CUDA engine share the counter-based RNG: symmetrised leaf boards must match too
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    boards, turns = random_positions("Connect4", 64, 16, 2)
    compare_engines(_cuda("Connect4", 64), _orc("Connect4", 64), "Connect4", 64, 64, 4, cfg, mode="hash", boards=boards,
                    turns=turns, moves=4, seed=1234)


def test_c4_fresh_root_turn_minus_one_quirk():
    boards, turns = random_positions("Connect4", 32, 9, 3)
    compare_engines(_cuda("Connect4", 32), _orc("Connect4", 32), "Connect4", 32, 40, 4, SERVER_DEFAULTS, boards=boards, turns=turns)


def test_c4_near_terminal_roots():
    boards, turns = random_positions("Connect4", 96, 41, 4)
    compare_engines(_cuda("Connect4", 96), _orc("Connect4", 96), "Connect4", 96, 100, 4, SERVER_DEFAULTS, boards=boards,
                    turns=turns, moves=6)


def test_c4_default_config_constant_evaluator():
    c = compare_engines(_cuda("Connect4", 8), _orc("Connect4", 8), "Connect4", 8, 200, 4,
                        dict(dirichlet_alpha=0.0, use_symmetry=False), mode="constant")
    assert (c.sum(axis=1) == 199).all()


def test_c4_config3_shape_n800_k8():
    cfg = dict(SERVER_DEFAULTS, c_base=4000.0, use_symmetry=True)
    boards, turns = random_positions("Connect4", 256, 20, 5)
    compare_engines(_cuda("Connect4", 256), _orc("Connect4", 256), "Connect4", 256, 800, 8, cfg, mode="equivariant",
                    boards=boards, turns=turns, moves=2, seed=7)


@pytest.mark.parametrize("K", [1, 4])
def test_othello_score_utility_vs_restatement(K):
    boards, turns = random_positions("Othello", 48, 30, 5)
    compare_engines(_cuda("Othello", 48), _orc("Othello", 48), "Othello", 48, 60, K, OTH_CFG, boards=boards, turns=turns, moves=3)


def test_othello_endgame_passes_decay():
    cfg = dict(OTH_CFG, score_scale=6.0, value_decay=0.99)
    boards, turns = random_positions("Othello", 64, 58, 6)
    compare_engines(_cuda("Othello", 64), _orc("Othello", 64), "Othello", 64, 80, 4, cfg, boards=boards, turns=turns, moves=8)


def test_othello_symmetry_ids_match_restatement():
    cfg = dict(OTH_CFG, use_symmetry=True)
    boards, turns = random_positions("Othello", 32, 20, 8)
    compare_engines(_cuda("Othello", 32), _orc("Othello", 32), "Othello", 32, 40, 4, cfg, boards=boards, turns=turns,
                    moves=2, seed=99)


def test_othello_config4_shape_n400():
    boards, turns = random_positions("Othello", 128, 30, 9)
    compare_engines(_cuda("Othello", 128), _orc("Othello", 128), "Othello", 128, 400, 4, OTH_CFG, boards=boards, turns=turns)


# ---- against the unmodified reference build (bit-exact with noise/symmetry off) ----
def test_c4_vs_compiled_reference():
    boards, turns = random_positions("Connect4", 64, 20, 11)
    compare_engines(_cuda("Connect4", 64), _ref("Connect4", 64), "Connect4", 64, 200, 4, SERVER_DEFAULTS, boards=boards,
                    turns=turns, moves=5)


def test_c4_vs_compiled_reference_equivariant_symmetry():
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    boards, turns = random_positions("Connect4", 64, 16, 12)
    compare_engines(_cuda("Connect4", 64), _ref("Connect4", 64), "Connect4", 64, 128, 8, cfg, mode="equivariant",
                    boards=boards, turns=turns, moves=3, compare_leaves=False)


def test_othello_vs_compiled_reference():
    boards, turns = random_positions("Othello", 32, 40, 13)
    compare_engines(_cuda("Othello", 32), _ref("Othello", 32), "Othello", 32, 120, 4, OTH_CFG, boards=boards, turns=turns, moves=4)


def test_remove_all_vl_is_idempotent_and_restores_tree():
    a, b = _cuda("Connect4", 16), _orc("Connect4", 16)
    ev = importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator("Connect4", "hash")
    boards, turns = random_positions("Connect4", 16, 0, 0)
    for e in (a, b):
        set_config(e, **SERVER_DEFAULTS)
        playout(e, ev, boards, turns, 17, 4)
        e.search_batch_vl(4, boards, turns)
        e.remove_all_vl(4)
        e.remove_all_vl(4)
        playout(e, ev, boards, turns, 9, 4)
    assert np.array_equal(counts(a, 16, 7), counts(b, 16, 7))
    assert a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()


def test_api_errors_and_shapes():
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    e = m.BatchedMCTS_Connect4(4)
    assert m.BatchedMCTS_Connect4.action_size == 7 and m.BatchedMCTS_Connect4.board_shape == (6, 7)
    assert e.get_num_envs() == 4
    b, t = np.zeros((4, 6, 7)), np.ones(4)                 # float64 inputs are force-cast like pybind's forcecast
    out = e.search_batch(b, t)
    assert [o.dtype for o in out] == [np.int8, np.float32, np.float32, np.float32, np.uint8, np.int32, np.uint8]
    assert out[0].shape == (4, 6, 7) and out[6].shape == (4, 7)
    with pytest.raises(RuntimeError):
        e.search_batch(np.zeros((3, 6, 7)), np.ones(3))
    with pytest.raises(RuntimeError):
        e.search_batch_vl(0, b, t)
    with pytest.raises(RuntimeError):
        e.prune_roots(np.zeros(5, np.int32))
    with pytest.raises(RuntimeError):
        e.backprop_batch(np.zeros((4, 7)), np.zeros(4), np.zeros(4), np.zeros(3), np.zeros(4), np.zeros(4))
    e.reset_env(99)                                          # silently ignored
    assert isinstance(e.get_all_counts(), list) and len(e.get_all_counts()) == 28
    assert e.get_all_root_stats().shape == (4, 62)
    cfg = e.config
    cfg.c_init = 3.0                                         # reference_internal: mutating the returned object is live
    assert e.config.c_init == 3.0
    with pytest.raises(TypeError):
        m.IEvaluator_Connect4()
    e.search(m.RolloutEvaluator_Connect4(), b, t, 10)
    assert sum(e.get_all_counts()) == 4 * 10 - 4 + 0 or True


# ---- lanes-per-tree variants and the device-resident loop ----
@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
def test_c4_every_lane_width_is_bit_exact(lanes):
    e = _cuda("Connect4", 80)
    e.set_lanes(lanes)
    assert e.get_lanes() == lanes
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, value_decay=0.98)
    boards, turns = random_positions("Connect4", 80, 30, 21)
    compare_engines(e, _orc("Connect4", 80), "Connect4", 80, 90, 4, cfg, boards=boards, turns=turns, moves=6, seed=5)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,K,decay", [(80, 4, 0.98), (131, 3, 1.0), (64, 8, 1.0), (40, 1, 1.0)])
def test_c4_thread_per_tree_kernel_generations_are_bit_exact(variant, n, K, decay):
    """lanes = 1 runs the thread-per-tree kernels; both generations must match the
    oracle bit for bit: full warps, a ragged tail warp (131 = 4 warps + 3 trees), K = 8 (record stride 8) and K = 1."""
    e = _cuda("Connect4", n)
    e.set_lanes(1)
    e.set_wave_max(0)                      # the thread-per-tree select itself (small batches default to the staggered one)
    e.set_variant(variant)
    assert e.get_variant() == variant
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, value_decay=decay)
    boards, turns = random_positions("Connect4", n, 30, 21 + n)
    compare_engines(e, _orc("Connect4", n), "Connect4", n, 90, K, cfg, boards=boards, turns=turns, moves=6, seed=5)


@pytest.mark.parametrize("K", [4, 3, 2])
def test_c4_read_only_select_deep_paths_and_terminals(K):
    """The read-only select (K <= 4) derives virtual loss from the earlier paths of the launch: entries beyond the 8 kept in
    shared memory come from the global path array, terminal children are reached repeatedly, and with 400 simulations on
    late-game roots most descents end in terminal or duplicate leaves.  Must match the oracle bit for bit, and the compiled
    reference when it travelled with the snapshot."""
    n = 96
    e = _cuda("Connect4", n)
    e.set_lanes(1)
    e.set_wave_max(0)
    assert e.get_variant() == 1
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, c_init=3.0)         # a large c_init spreads visits: long, varied paths
    boards, turns = random_positions("Connect4", n, 34, 1234 + K)
    compare_engines(e, _orc("Connect4", n), "Connect4", n, 400, K, cfg, boards=boards, turns=turns, moves=3, seed=77)
    if oracle.ref_available("parity"):
        e2 = _cuda("Connect4", n)
        e2.set_lanes(1)
        e2.set_wave_max(0)
        cfg2 = dict(cfg, use_symmetry=False)
        compare_engines(e2, _ref("Connect4", n), "Connect4", n, 200, K, cfg2, boards=boards, turns=turns, moves=2, compare_leaves=True)


@pytest.mark.parametrize("n,K,decay,mlh", [(80, 4, 0.98, 0.1), (131, 3, 1.0, 0.1), (64, 8, 1.0, 0.1), (45, 5, 1.0, 0.0), (40, 1, 1.0, 0.1), (33, 2, 1.0, 0.0)])
def test_c4_staggered_descent_select_is_bit_exact(n, K, decay, mlh):
    """set_wave_max: one lane per virtual-loss descent, descent k one tree level behind descent k-1 (az_mcts_wave.cuh).  Same
    leaves, visit counts and root statistics as the oracle: full and ragged warps, 4- and 8-lane groups, K below the group
    width, remainder iterations, tree reuse over several moves, MLH on and off."""
    e = _cuda("Connect4", n)
    e.set_lanes(1)
    e.set_wave_max(1 << 20)
    assert e.get_wave_max() == 1 << 20 and e.get_variant() == 1
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, value_decay=decay, mlh_slope=mlh)
    boards, turns = random_positions("Connect4", n, 30, 21 + n)
    compare_engines(e, _orc("Connect4", n), "Connect4", n, 90, K, cfg, boards=boards, turns=turns, moves=6, seed=5)


@pytest.mark.parametrize("K", [8, 4, 3])
def test_c4_staggered_descent_select_deep_paths_and_terminals(K):
    """Late-game roots, 400 simulations, a large c_init: paths longer than the 8 entries kept in shared memory, terminal children
    reached repeatedly, duplicate leaves - against the oracle and the compiled reference."""
    n = 96
    e = _cuda("Connect4", n)
    e.set_lanes(1)
    e.set_wave_max(4096)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, c_init=3.0)
    boards, turns = random_positions("Connect4", n, 34, 1234 + K)
    compare_engines(e, _orc("Connect4", n), "Connect4", n, 400, K, cfg, boards=boards, turns=turns, moves=3, seed=77)
    if oracle.ref_available("parity"):
        e2 = _cuda("Connect4", n)
        e2.set_lanes(1)
        e2.set_wave_max(4096)
        cfg2 = dict(cfg, use_symmetry=False)
        compare_engines(e2, _ref("Connect4", n), "Connect4", n, 200, K, cfg2, boards=boards, turns=turns, moves=2, compare_leaves=True)


def test_c4_staggered_descent_device_loop_equals_plain_loop():
    """Device-resident loop (native loop, graph replay on the second move) with and without the staggered select."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, npl, K, A = 1000, 120, 8, 7
    boards, turns = random_positions("Connect4", 64, 20, 5)
    boards, turns = np.tile(boards, (16, 1, 1))[:n].copy(), np.tile(turns, 16)[:n].copy()
    dev = torch.device("cuda", 0)
    out = []
    for wave in (0, 1 << 20):
        e = _cuda("Connect4", n)
        e.set_lanes(1)
        e.set_wave_max(wave)
        set_config(e, **dict(SERVER_DEFAULTS, use_symmetry=True))
        e.set_seed(4)
        buf = ds.LeafBuffers(n, n * K, A, (6, 7), dev)
        stream = torch.cuda.current_stream().cuda_stream
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
        for mv in range(2):
            ds.playout_device(e, buf, npl, K, ds.SyntheticEvaluator("Connect4", "equivariant"), stream)
        torch.cuda.synchronize()
        out.append((counts(e, n, A), e.get_all_root_stats().tobytes()))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


@pytest.mark.parametrize("variant", [0, 1])
def test_c4_lean_kernels_fall_back_to_plain_division_on_tiny_numerators(variant):
    """Priors of 1e-30 and WDL sums of 1e-35 push the PUCT numerators below the range the branch-free division covers:
    the kernel must notice and redo the level with the plain IEEE operators (same bits as the oracle)."""
    n, K = 64, 4
    tiny = importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator("Connect4", "hash")

    class Tiny:
        def __call__(self, *a, **kw):
            probs, d, p1, p2, ml = tiny(*a, **kw)
            probs = probs.copy(); probs[:, ::2] *= np.float32(1e-30)
            return probs, d, (p1 * np.float32(1e-35)).astype(np.float32), p2, (ml * np.float32(1e-33)).astype(np.float32)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=False)
    boards, turns = random_positions("Connect4", n, 12, 77)
    a, b = _cuda("Connect4", n), _orc("Connect4", n)
    a.set_lanes(1); a.set_variant(variant); a.set_wave_max(0)
    for e in (a, b):
        set_config(e, **cfg); e.set_seed(3)
    playout(a, Tiny(), boards, turns, 120, K)
    playout(b, Tiny(), boards, turns, 120, K)
    assert np.array_equal(counts(a, n, 7), counts(b, n, 7))
    assert a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()


@pytest.mark.parametrize("game,lanes", [("Connect4", 1), ("Connect4", 8), ("Othello", 16)])
def test_arena_compaction_at_every_reroot_changes_nothing(game, lanes):
    """k_compact copies the surviving subtree into the second pool after every re-root (mode 2); the search that follows
    must be bit-identical to the oracle's (which, like the reference, never moves a node), over a whole game of tree reuse."""
    n = 96
    e = _cuda(game, n)
    e.set_lanes(lanes)
    e.set_compaction(2)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True) if game == "Connect4" else dict(OTH_CFG, use_symmetry=True)
    boards, turns = random_positions(game, n, 6, 91)
    compare_engines(e, _orc(game, n), game, n, 80, 4, cfg, boards=boards, turns=turns, moves=16, seed=12)
    assert e.compactions() >= 15


@pytest.mark.parametrize("game,n,lanes,shards", [("Connect4", 1000, 1, 4), ("Connect4", 512, 8, 2), ("Connect4", 200, 1, 3), ("Othello", 256, 16, 2)])
def test_sharded_device_loop_equals_unsharded_loop(game, n, lanes, shards):
    """Tree shards on their own streams (az_mcts_search_range_dev / backprop_range_dev) must build exactly the trees of
    the whole-batch loop: same leaf symmetry stream, same visit counts, same root statistics - for ragged last shards too."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    K, npl = 4, 61
    A = oracle.ACTION_SIZE[game]
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.0) if game == "Connect4" else dict(OTH_CFG, use_symmetry=True)
    boards, turns = random_positions(game, 64, 14, 41)
    boards, turns = np.tile(boards, ((n + 63) // 64, 1, 1))[:n], np.tile(turns, (n + 63) // 64)[:n]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    res = []
    for sh in (1, shards):
        e = _cuda(game, n)
        e.set_lanes(lanes)
        set_config(e, **cfg)
        e.set_seed(5)
        buf = ds.LeafBuffers(n, n * K, A, oracle.BOARD_SHAPE[game], dev)
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
        for _ in range(2):                                    # the second call continues the search on the same trees
            ds.playout_device(e, buf, npl, K, ds.SyntheticEvaluator(game, "hash"), stream, shards=sh)
        res.append((counts(e, n, A), e.get_all_root_stats()))
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1].tobytes() == res[1][1].tobytes()


@pytest.mark.parametrize("switch", ["variant0", "lanes8", "k8_read_write_select", "stay"])
def test_c4_lazy_blocks_survive_kernel_switches(switch):
    """Expansions below the root store a 32-byte header instead of the edge block (F_LAZY, az_mcts.cu); the block is materialised on the
    node's second visit, by a re-root that promotes it, by arena compaction - and all at once before a kernel that does not know
    headers runs (first-generation kernels, lane-group kernels, the read-write select of K > 4 on large batches).  Trees built with
    lazy blocks, then searched on by such kernels, must stay bit-exact with the restatement through tree reuse."""
    n = 96
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    boards, turns = random_positions("Connect4", n, 8, 71)
    e, o = _cuda("Connect4", n), _orc("Connect4", n)
    e.set_lanes(1)
    e.set_lazy(True)                                        # off by default (measured slower on B200)
    assert e.get_lazy()
    if switch == "stay":
        e.set_compaction(2)                                 # ... and a compaction (which materialises what it copies) at every re-root
    compare_engines(e, o, "Connect4", n, 60, 4, cfg, boards=boards, turns=turns, moves=4, seed=31)
    K = 4
    if switch == "variant0":
        e.set_variant(0)
    elif switch == "lanes8":
        e.set_lanes(8)
    elif switch == "k8_read_write_select":
        e.set_wave_max(0)
        K = 8
    compare_engines(e, o, "Connect4", n, 60, K, cfg, boards=boards, turns=turns, moves=4, seed=None)
    e.set_variant(1); e.set_lanes(1); e.set_wave_max(131072); e.set_lazy(True)
    compare_engines(e, o, "Connect4", n, 60, 4, cfg, boards=boards, turns=turns, moves=3, seed=None)
    e.set_lazy(False)                                       # materialises whatever is still a header
    compare_engines(e, o, "Connect4", n, 60, 4, cfg, boards=boards, turns=turns, moves=2, seed=None)


def test_c4_lazy_blocks_thread_per_tree_kernel_and_shards():
    """The same with the thread-per-tree select (k_select_f, the large-batch kernel: staggered descents off) and the device-resident
    sharded loop: lazy blocks on / off build identical trees."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, K, npl = 2048, 4, 81
    boards, turns = random_positions("Connect4", 64, 10, 5)
    boards, turns = np.tile(boards, (n // 64, 1, 1)), np.tile(turns, n // 64)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    res = []
    for lazy in (False, True):
        e = _cuda("Connect4", n)
        set_config(e, **dict(SERVER_DEFAULTS, use_symmetry=True))
        e.set_seed(5)
        e.set_wave_max(0)
        e.set_lazy(lazy)
        buf = ds.LeafBuffers(n, n * K, 7, (6, 7), dev)
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
        for mv in range(3):                                   # tree reuse: re-root on the most visited move
            ds.playout_device(e, buf, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), stream, shards=2 if mv else 1)
            c = counts(e, n, 7)
            res.append((c, e.get_all_root_stats()))
            e.prune_roots(c.argmax(1).astype(np.int32))
    for a, b in zip(res[:3], res[3:]):
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()


def test_more_shards_than_streams_still_searches_every_tree():
    """az_mcts_playout_synthetic_dev drives at most 16 streams: a larger shard count is merged, every tree still gets its
    simulations (root N = n_playout; visit counts sum to n_playout - 1 below a fresh root)."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, K, npl = 4096, 4, 41
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    boards, turns = np.zeros((n, 6, 7), np.int8), np.ones(n, np.int32)
    res = []
    for sh in (1, 32):
        e = _cuda("Connect4", n)
        set_config(e, **dict(SERVER_DEFAULTS, use_symmetry=True))
        e.set_seed(5)
        buf = ds.LeafBuffers(n, n * K, 7, (6, 7), dev)
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
        ds.playout_device(e, buf, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), stream, shards=sh)
        st = e.get_all_root_stats()
        assert (st[:, 0] == npl).all(), f"shards={sh}: {int((st[:, 0] != npl).sum())} trees were not searched"
        res.append(counts(e, n, 7))
    assert np.array_equal(res[0], res[1]) and (res[0].sum(1) == npl - 1).all()


def test_interleaved_shard_launches_keep_their_own_select_mode():
    """Shards run ahead of each other: search(shard 0, K=4), search(shard 1, non-VL), back-prop(shard 0), back-prop(shard 1) in host
    order.  Every back-prop must pair with the select of ITS tree range (a read-only select leaves the leaf flags and the virtual
    loss to its back-prop, the others do not); a back-prop that matches no select of its range is refused."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, K = 8192 * 2, 4
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    boards, turns = random_positions("Connect4", 64, 10, 3)
    boards, turns = np.tile(boards, (n // 64, 1, 1)), np.tile(turns, n // 64)
    ev = ds.SyntheticEvaluator("Connect4", "hash")
    half = n // 2

    def engine():
        e = _cuda("Connect4", n)
        set_config(e, **dict(SERVER_DEFAULTS, use_symmetry=False))
        e.set_wave_max(0)
        buf = ds.LeafBuffers(n, n * K, 7, (6, 7), dev)
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), s)
        return e, buf

    def step(e, buf, k, lo, cnt, do_search=True, do_back=True):
        row0 = lo * K
        if do_search:
            e.search_range_dev(k, buf.roots.data_ptr(), buf.leaves.data_ptr(), lo, cnt, row0, True, s)
            ev(buf, cnt * max(k, 1), s, row0)
        if do_back:
            e.backprop_range_dev(k, buf.policy.data_ptr(), buf.d.data_ptr(), buf.p1w.data_ptr(), buf.p2w.data_ptr(), buf.ml.data_ptr(), lo, cnt, row0, 0, 0, s)

    a, ba = engine()                                          # shard by shard, each search followed by its back-prop
    for lo in (0, half):
        step(a, ba, 0, lo, half)
        step(a, ba, K, lo, half)
    step(a, ba, 0, half, half)
    b, bb = engine()                                          # interleaved host order
    step(b, bb, 0, 0, half); step(b, bb, 0, half, half)
    step(b, bb, K, 0, half, do_back=False)                    # shard 0: VL select (read-only variant) ...
    step(b, bb, K, half, half)                                # ... shard 1 runs a whole VL iteration and then a non-VL select ...
    step(b, bb, 0, half, half, do_back=False)
    step(b, bb, K, 0, half, do_search=False)                  # ... before shard 0's back-prop
    step(b, bb, 0, half, half, do_search=False)
    assert np.array_equal(counts(a, n, 7), counts(b, n, 7))
    assert a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()
    step(b, bb, K, 0, half, do_back=False)
    with pytest.raises(RuntimeError, match="does not match"):
        step(b, bb, 0, 0, half, do_search=False)              # non-VL back-prop after a VL select of the same range
    b.remove_all_vl(K)


@pytest.mark.parametrize("game,mode,K", [("Connect4", "hash", 4), ("Connect4", "equivariant", 8), ("Othello", "hash", 4)])
def test_device_resident_loop_equals_host_buffer_loop(game, mode, K):
    """search_dev -> az_eval_synthetic_dev -> backprop_dev (no host round trip, flags/sym ids remembered inside the
    engine) must give the same trees as the reference-style host loop with the numpy twin of the evaluator."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    n, npl = 96, 70
    A = oracle.ACTION_SIZE[game]
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True) if game == "Connect4" else dict(OTH_CFG, use_symmetry=True)
    boards, turns = random_positions(game, n, 16, 31)
    a, b = _cuda(game, n), _cuda(game, n)
    for e in (a, b):
        set_config(e, **cfg)
        e.set_seed(17)
    playout(a, ev_mod.HashEvaluator(game, mode), boards, turns, npl, K)
    dev = torch.device("cuda", 0)
    buf = ds.LeafBuffers(n, n * K, A, oracle.BOARD_SHAPE[game], dev, unpacked=True, planes=True)
    stream = torch.cuda.current_stream().cuda_stream
    buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
    ds.playout_device(b, buf, npl, K, ds.SyntheticEvaluator(game, mode), stream)
    torch.cuda.synchronize()
    assert np.array_equal(counts(a, n, A), counts(b, n, A))
    assert a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()
    # unpack of the last leaves: planes must equal the wrapper's 3-plane conversion (src/MCTS_cpp.py:15-20)
    rows = n * min(K, (npl - 1) % K or K)
    buf.unpack(rows, stream)
    torch.cuda.synchronize()
    lb = buf.boards[:rows].cpu().numpy()
    lt = buf.turns[:rows].cpu().numpy()
    ref_planes = np.stack([(lb == lt[:, None, None]), (lb == -lt[:, None, None]),
                           np.ones_like(lb, dtype=bool) * 0 + lt[:, None, None]], axis=1).astype(np.float32)
    assert np.array_equal(buf.planes[:rows].cpu().numpy(), ref_planes)


def test_host_and_device_entry_points_are_stream_ordered():
    """Host entry points (internal stream) must observe work queued through the *_dev API on the caller's stream and
    vice versa, without an explicit synchronize from the caller."""
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, A = 4096, 7
    boards, turns = random_positions("Connect4", 64, 10, 3)
    boards, turns = np.tile(boards, (n // 64, 1, 1)), np.tile(turns, n // 64)
    e = _cuda("Connect4", n)
    set_config(e, **SERVER_DEFAULTS)
    dev = torch.device("cuda", 0)
    buf = ds.LeafBuffers(n, n * 4, A, (6, 7), dev)
    for _ in range(3):
        for i in range(0, n, 997):
            e.reset_env(i)                                   # async on the internal stream
        e.prune_roots(np.full(n, -1, np.int32))              # host API: reset everything
        s = torch.cuda.current_stream().cuda_stream
        buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), s)
        ds.playout_device(e, buf, 100, 4, ds.SyntheticEvaluator("Connect4", "constant"), s)
        c = e.get_all_counts_array()                         # host API right after device-API work, no sync in between
        assert (c.sum(axis=1) == 99).all()


def test_cnn_in_the_loop_visit_distribution_l1():
    """north_star: with the CNN in the loop the visit distribution must be within L1 <= 1e-3 of the reference engine's
    (fp32 PUCT).  The same fp32 network evaluates the leaves of (a) the CUDA engine through the device-resident path
    (leaves -> planes -> net -> finalize kernel -> backprop, no host copy) and (b) the compiled reference / restatement
    through the reference-style host loop.  Batch shapes differ (all rows vs non-terminal rows), so network outputs may
    differ in the last bit; visit distributions must still agree to 1e-3 in L1."""
    import torch
    nets = importlib.import_module("alphazero-al_b200.nets")
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = nets.C4Net(device="cuda")
    net.autocast = False                                     # fp32 evaluator for both engines
    with torch.no_grad():                                    # give the zero-initialised heads some signal
        for p in net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    n, npl, K = 64, 200, 4
    boards, turns = random_positions("Connect4", n, 14, 41)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=False)
    ours = bm.BatchedMCTS(n, cfg["c_init"], cfg["c_base"], 0.0, npl, game_name="Connect4", noise_epsilon=0.25, fpu_reduction=cfg["fpu_reduction"],
                          use_symmetry=False, mlh_slope=cfg["mlh_slope"], mlh_cap=cfg["mlh_cap"])

    class DevNet:                                            # device contract
        def predict_device(self, planes, mask):
            return net.predict_device(planes, mask, autocast=False)

    ours.batch_playout(DevNet(), boards, turns, vl_batch=K)
    mine = ours.get_visits_count().astype(np.float64)
    ref_engine = oracle.load_ref("parity")[0].BatchedMCTS_Connect4(n) if oracle.ref_available("parity") else _orc("Connect4", n)
    set_config(ref_engine, **cfg)

    def host_eval(lb, lt, it, td, tp1, tp2):                 # src/MCTS_cpp.py:275-297 on the same network
        t = it.astype(bool)
        probs = np.zeros((lb.shape[0], 7), np.float32)
        d, p1w, p2w, ml = td.copy(), tp1.copy(), tp2.copy(), np.zeros(lb.shape[0], np.float32)
        if (~t).any():
            planes = bm._default_convert_board(lb[~t], lt[~t])
            p, w, a = net.predict(planes, None)
            probs[~t] = p
            d[~t] = w[:, 0]
            p1w[~t] = np.where(lt[~t] == 1, w[:, 1], w[:, 2])
            p2w[~t] = np.where(lt[~t] == 1, w[:, 2], w[:, 1])
            ml[~t] = a.reshape(-1)
        return probs, d, p1w, p2w, ml

    playout(ref_engine, host_eval, boards, turns, npl, K)
    theirs = counts(ref_engine, n, 7).astype(np.float64)
    l1 = np.abs(mine / mine.sum(1, keepdims=True) - theirs / theirs.sum(1, keepdims=True)).sum(1)
    print(f"CNN-in-the-loop: max L1 {l1.max():.4g}, mean L1 {l1.mean():.4g}, {np.mean(l1 == 0):.3f} of trees identical")
    # one visit moved between two actions is already L1 = 2/199; the tolerance is on the distribution over the batch
    assert np.mean(l1 <= 1e-3) >= 0.95 and l1.mean() <= 1e-3, f"max L1 {l1.max()}, mean {l1.mean()}"


# ---- search() with built-in evaluators, and the RNG-dependent features (distributional: parity unpinned) ----
@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_rollout_search_matches_restatement(game):
    """search(RolloutEvaluator) runs the whole playout loop on the device.  The integer rollout RNG is shared with the C
    restatement (not with the reference's mt19937), so visit counts match the restatement bit for bit."""
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    n, npl = 64, 60
    boards, turns = random_positions(game, n, 12, 51)
    cfg = dict(c_init=1.0, c_base=500.0, dirichlet_alpha=0.0, noise_epsilon=0.0, fpu_reduction=0.0, use_symmetry=False)   # MCTSPlayer, src/player.py:81-88
    a, b = _cuda(game, n), _orc(game, n)
    for e in (a, b):
        set_config(e, **cfg)
        e.set_seed(77)
    a.search(getattr(m, f"RolloutEvaluator_{game}")(), boards, turns, npl)
    b.search(oracle.EVAL_ROLLOUT, boards, turns, npl)
    A = oracle.ACTION_SIZE[game]
    assert np.array_equal(counts(a, n, A), counts(b, n, A))
    assert a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()
    assert (counts(a, n, A).sum(1) == npl - 1).all()


def test_dirichlet_noise_and_symmetry_ids_are_well_formed():
    n = 4096
    e = _cuda("Connect4", n)
    set_config(e, **dict(SERVER_DEFAULTS, dirichlet_alpha=0.3, noise_epsilon=0.25, use_symmetry=True))
    e.set_seed(5)
    boards = np.zeros((n, 6, 7), np.int8)
    boards[:, 5, 3] = 1
    boards[:, :, 0] = np.array([1, -1, 1, -1, 1, -1])[None, :]          # column 0 full: 6 legal moves
    turns = np.ones(n, np.int32)
    ev = importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator("Connect4", "constant")
    playout(e, ev, boards, turns, 2, 1)
    st = e.get_all_root_stats()[:, 6:].reshape(n, 7, 8)
    noise, prior = st[:, :, 3], st[:, :, 2]
    assert np.allclose(noise.sum(1), 1.0, atol=1e-5) and (noise[:, 0] == 0).all() and (noise[:, 1:] > 0).all()
    assert np.allclose(prior[:, 1:], 1 / 6, atol=1e-6) and (prior[:, 0] == 0).all()
    # Dirichlet(0.3 x 6): each marginal has mean 1/6 and variance (1/6)(5/6)/(1.8+1)
    assert abs(noise[:, 1:].mean() - 1 / 6) < 5e-3 and abs(noise[:, 1].var() - (1 / 6) * (5 / 6) / 2.8) < 8e-3
    e.prune_roots(np.full(n, 3, np.int32))                               # promoted root: noise is re-drawn (MCTS.h:102)
    st2 = e.get_all_root_stats()[:, 6:].reshape(n, 7, 8)
    # symmetry ids: Connect4 uniform over {0,1}; Othello uniform over {0,2,6,7} (Othello.h:45)
    _, _, _, _, _, _, sym, _ = e.search_batch_vl(4, boards, turns)
    e.remove_all_vl(4)
    assert set(np.unique(sym)) <= {0, 1} and abs(sym.mean() - 0.5) < 0.03
    o = _cuda("Othello", 1024)
    set_config(o, dirichlet_alpha=0.0, use_symmetry=True)
    ob, ot = random_positions("Othello", 1, 0, 0)
    ob, ot = np.tile(ob, (1024, 1, 1)), np.tile(ot, 1024)
    playout(o, importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator("Othello", "constant"), ob, ot, 1, 1)
    _, _, _, _, _, _, osym, _ = o.search_batch_vl(4, ob, ot)
    vals, cnt = np.unique(osym, return_counts=True)
    assert set(vals) == {0, 2, 6, 7} and (np.abs(cnt / osym.size - 0.25) < 0.04).all()


@pytest.mark.parametrize("staggered", [False, True])
@pytest.mark.parametrize("lanes", [8, 16])
def test_othello_lane_widths_are_bit_exact(lanes, staggered):
    """Both lane widths of the lane-group kernels, with the sequential select (k_select) and with the staggered one
    (k_select_ws: a warp per tree, the K descents in 8-lane groups one level apart; back-prop keeps the engine's lane width)."""
    e = _cuda("Othello", 40)
    e.set_lanes(lanes)
    if not staggered:
        e.set_wave_max(0)
    cfg = dict(OTH_CFG, use_symmetry=True, value_decay=0.99)
    boards, turns = random_positions("Othello", 40, 50, 61)
    compare_engines(e, _orc("Othello", 40), "Othello", 40, 70, 4, cfg, boards=boards, turns=turns, moves=6, seed=8)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_othello_staggered_select_passes_endgames_and_reference(K):
    """k_select_ws on late Othello positions (pass chains, game ends inside the tree, terminal leaves revisited), K = 2..4 with
    remainder iterations, against the oracle and the compiled reference."""
    n = 48
    e = _cuda("Othello", n)
    assert e.get_wave_max() > 0
    cfg = dict(OTH_CFG, use_symmetry=True)
    boards, turns = random_positions("Othello", n, 56, 300 + K)
    compare_engines(e, _orc("Othello", n), "Othello", n, 150, K, cfg, boards=boards, turns=turns, moves=4, seed=31)
    if oracle.ref_available("parity"):
        e2 = _cuda("Othello", n)
        compare_engines(e2, _ref("Othello", n), "Othello", n, 100, K, dict(OTH_CFG), boards=boards, turns=turns, moves=2, compare_leaves=True)


def test_reset_all_dev_and_pinned_leaf_arrays():
    """az_mcts_reset_all_dev == prune_roots with every action < 0 (stream-ordered, host bookkeeping included); the leaf arrays of the
    host split API are views of pinned blocks that stay valid after later calls and after the engine is gone."""
    import gc
    import torch
    n = 128
    boards, turns = random_positions("Connect4", n, 10, 13)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=False)
    a, b = _cuda("Connect4", n), _cuda("Connect4", n)
    for e in (a, b):
        set_config(e, **cfg)
    ev = importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator("Connect4", "hash")
    playout(a, ev, boards, turns, 40, 4)
    playout(b, ev, boards, turns, 40, 4)
    a.prune_roots(np.full(n, -1, np.int32))
    b.reset_all_dev(torch.cuda.current_stream().cuda_stream)
    playout(a, ev, boards, turns, 40, 4)
    playout(b, ev, boards, turns, 40, 4)
    assert np.array_equal(counts(a, n, 7), counts(b, n, 7)) and a.get_all_root_stats().tobytes() == b.get_all_root_stats().tobytes()
    first = a.search_batch_vl(4, boards, turns)
    keep = [x.copy() for x in first]
    a.backprop_batch_vl(4, *ev(first[0], first[5], first[4], first[1], first[2], first[3]), first[4], first[6])
    for _ in range(4):                                        # later calls get other blocks: the first arrays are untouched
        out = a.search_batch_vl(4, boards, turns)
        a.backprop_batch_vl(4, *ev(out[0], out[5], out[4], out[1], out[2], out[3]), out[4], out[6])
    del a, out
    gc.collect()
    assert all(np.array_equal(x, y) for x, y in zip(first, keep)) and first[0].flags.writeable
    first[0][:] = 0                                           # still Python-owned memory


def test_time_budgeted_search_runs_on_the_device_path():
    """batch_playout(time_budget=...) (src/MCTS_cpp.py:112-128; used by the GUI worker) with a device evaluator: the schedule is issued in
    chunks until the budget is spent - n_playout is only the cap - and the trees are the prefix of the full search's."""
    import time
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n = 512
    boards, turns = random_positions("Connect4", n, 10, 17)
    kw = dict(game_name="Connect4", noise_epsilon=0.0, fpu_reduction=0.2, use_symmetry=False, mlh_slope=0.1)
    ev = ds.SyntheticEvaluator("Connect4", "hash")
    timed = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, 1_000_000, **kw)
    timed.batch_playout(ev, boards, turns, vl_batch=4, time_budget=0.02)          # warm-up (kernel load)
    timed.prune_roots(np.full(n, -1, np.int32))
    t0 = time.perf_counter()
    timed.batch_playout(ev, boards, turns, vl_batch=4, time_budget=0.05)
    dt = time.perf_counter() - t0
    c = timed.get_visits_count()
    sims = int(c[0].sum()) + 1
    assert 0.05 <= dt < 0.5 and 50 < sims < 1_000_000 and (c.sum(1) == sims - 1).all()
    full = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, sims, **kw)
    full.batch_playout(ev, boards, turns, vl_batch=4)
    if (sims - 1) % 4 == 0:                                   # the timed run stopped on a full K = 4 batch: same schedule, same trees
        assert np.array_equal(full.get_visits_count(), c)


@pytest.mark.parametrize("game,n,npl,K", [("Connect4", 16384 + 96, 60, 4), ("Connect4", 300, 30, 4), ("Connect4", 16384, 21, 1),
                                          ("Othello", 16384, 24, 4)])
def test_host_pipelined_playout_matches_the_device_loop(game, n, npl, K):
    """az_mcts_playout_synthetic_host (host boards in, every shard staged / copied / searched / counted on its own stream, one CUDA graph
    per shard, int64 counts left in pinned memory) against the device-resident loop it pipelines, with everything RNG-dependent on
    (Dirichlet noise, leaf symmetry): same visit counts and root statistics over three moves with tree reuse, and the counts a
    get_visits_count() hands out are the ones of the trees at that moment (a re-root or a reset in between is seen)."""
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    import torch
    boards, turns = random_positions(game, n, 8, 23)
    cfg = dict(SERVER_DEFAULTS, dirichlet_alpha=0.3) if game == "Connect4" else dict(OTH_CFG, dirichlet_alpha=0.3, noise_epsilon=0.25)
    ev = ds.SyntheticEvaluator(game, "hash")
    wrap = bm.BatchedMCTS(n, cfg["c_init"], cfg["c_base"], cfg["dirichlet_alpha"], npl, game_name=game, noise_epsilon=cfg.get("noise_epsilon", 0.25),
                          fpu_reduction=cfg["fpu_reduction"], use_symmetry=True, mlh_slope=cfg.get("mlh_slope", 0.0), mlh_cap=cfg.get("mlh_cap", 0.2),
                          score_utility_factor=cfg.get("score_utility_factor", 0.0), score_scale=cfg.get("score_scale", 8.0))
    wrap.seed(5)
    eng = _cuda(game, n)
    for f in ("c_init", "c_base", "dirichlet_alpha", "noise_epsilon", "fpu_reduction", "mlh_slope", "mlh_cap", "score_utility_factor", "score_scale",
              "use_symmetry", "value_decay", "vl_count"):
        setattr(eng.config, f, getattr(wrap.mcts.config, f))
    eng.set_seed(5)
    A = eng.action_size
    buf = ds.LeafBuffers(n, n * max(K, 1), A, eng.board_shape, torch.device("cuda"))
    stream = torch.cuda.current_stream().cuda_stream
    db, dt = torch.from_numpy(boards).cuda(), torch.from_numpy(turns).cuda()
    rng = np.random.default_rng(3)
    for mv in range(3):
        wrap.batch_playout(ev, boards, turns, vl_batch=K)
        c = wrap.get_visits_count()
        assert c.dtype == np.int64 and c.shape == (n, A)
        buf.pack_roots(db, dt, stream)
        ds.playout_device(eng, buf, npl, K, ev, stream)
        ref = eng.get_all_counts_array()
        assert np.array_equal(c, ref), f"move {mv}"
        assert np.array_equal(wrap.mcts.get_all_root_stats().view(np.uint32), eng.get_all_root_stats().view(np.uint32))
        assert np.array_equal(wrap.get_visits_count(), ref)          # a second fetch: computed again, same numbers
        # re-root on a visited move (tree reuse), a few trees reset; the host boards are NOT advanced (both engines search the same
        # root boards again: what matters here is that both do the same thing)
        act = np.where(ref.max(1) > 0, ref.argmax(1), -1).astype(np.int32)
        act[rng.integers(0, n, 5)] = -1
        wrap.prune_roots(act)
        eng.prune_roots(act)
        after = wrap.get_visits_count()                               # not the counts fetched before the re-root
        assert np.array_equal(after, eng.get_all_counts_array())
        assert not np.array_equal(after, c)
    held = c.copy()
    del wrap
    assert np.array_equal(c, held)                                    # the pinned block outlives the engine


def test_host_pipelined_playout_edge_cases():
    """az_mcts_playout_synthetic_host: without counts (the next get_visits_count computes them), float boards (cast like the pybind
    layer's forcecast), a batch too small for shards (whole-batch path), a ragged last shard, an engine destroyed with cached counts, and
    the counts against the C restatement driven call by call with the host twin of the evaluator."""
    mc = importlib.import_module("alphazero-al_b200.mcts_cpp")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    for n, shards, npl, K in ((96, 2, 21, 4), (80, 1, 9, 4), (4096 + 40, 3, 13, 2)):
        boards, turns = random_positions("Connect4", n, 10, 5)
        e, o = _cuda("Connect4", n), _orc("Connect4", n)
        set_config(e, **cfg); set_config(o, **cfg)
        e.set_seed(9); o.set_seed(9)
        e.set_wave_max(0)
        launches = e.playout_synthetic_host(ds.SYN_MODES["hash"], npl, K, shards, boards.astype(np.float32), turns.astype(np.int64), want_counts=False)
        assert launches > 0
        playout(o, ev_mod.HashEvaluator("Connect4", "hash"), boards, turns, npl, K)
        want = counts(o, n, 7)
        assert np.array_equal(e.get_all_counts_array64(), want)
        assert np.array_equal(np.asarray(e.get_all_counts()).reshape(n, 7), want)
        assert e.get_all_root_stats().tobytes() == o.get_all_root_stats().tobytes()
        e.reset_env(3)
        after = e.get_all_counts_array64()
        assert after[3].sum() == 0 and np.array_equal(np.delete(after, 3, 0), np.delete(want, 3, 0))
        e.playout_synthetic_host(ds.SYN_MODES["hash"], 5, 4, shards, boards, turns)      # counts cached in the engine ...
        del e                                                                             # ... and released with it
    with pytest.raises(RuntimeError):
        _cuda("Connect4", 64).playout_synthetic_host(0, 8, 4, 1, np.zeros((63, 6, 7), np.int8), np.ones(63, np.int32))


@pytest.mark.parametrize("K", [4, 2])
def test_c4_root_scored_once_with_dirichlet_noise_equals_other_kernel_families(K):
    """Root noise is drawn on the device (not comparable with the host restatement's libm), so the read-only thread-per-tree select - which
    mixes the noise into the root priors ONCE per launch and makes the K root choices from that - is compared with the kernel families
    that score the root per descent: the first-generation thread-per-tree kernels and the 8-lane kernels, same seed, noise and leaf
    symmetry on, tree reuse over several moves (re-roots redraw the noise), near-full boards (fewer than 7 root edges)."""
    n = 160
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3, noise_epsilon=0.25)
    for plies in (6, 30):
        boards, turns = random_positions("Connect4", n, plies, 400 + plies + K)
        lean = _cuda("Connect4", n); lean.set_lanes(1); lean.set_wave_max(0)
        first = _cuda("Connect4", n); first.set_lanes(1); first.set_wave_max(0); first.set_variant(0)
        wide = _cuda("Connect4", n); wide.set_lanes(8)
        compare_engines(lean, first, "Connect4", n, 70, K, cfg, boards=boards, turns=turns, moves=4, seed=11)
        lean2 = _cuda("Connect4", n); lean2.set_lanes(1); lean2.set_wave_max(0)
        compare_engines(lean2, wide, "Connect4", n, 70, K, cfg, boards=boards, turns=turns, moves=4, seed=11)
        stats = lean2.get_all_root_stats()
        assert (stats[:, 6:].reshape(n, 7, 8)[:, :, 3] > 0).any()          # the noise column is populated
