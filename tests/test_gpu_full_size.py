"""Full-size runs of the BASELINE configurations on the device-resident loop, checked through size-independent
properties and against the C restatement on a random sample of the trees: trees are independent and the evaluators are pure
functions of the position, so tree i of an N-tree batch must equal, bit for bit, the same root searched in a small oracle
batch (visit counts and all 6+8A root statistics).  Covers what the small parity tests cannot: 32-bit arena addressing at
65 536 trees, shard boundaries of the real shard size, graph replay of the native loop, tail CTAs."""
import importlib
import os
import sys

import numpy as np
import pytest

import oracle
from harness import SERVER_DEFAULTS, counts, playout, random_positions, set_config

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

OTH_CFG = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False,
               score_utility_factor=0.15, score_scale=8.0)


def _c4_roots(n, seed):
    import bench
    return bench.c4_random_roots(n, seed)


def _oth_roots(n, seed):
    b, t = random_positions("Othello", 256, 30, seed)          # 256 distinct mid-game positions, tiled
    rep = (n + 255) // 256
    return np.tile(b, (rep, 1, 1))[:n].copy(), np.tile(t, rep)[:n].copy()


def _run(game, n, npl, K, cfg, mode, sample, shards, roots, moves=1):
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    mc = importlib.import_module("alphazero-al_b200.mcts_cpp")
    A = oracle.ACTION_SIZE[game]
    boards, turns = roots
    eng = getattr(mc, f"BatchedMCTS_{game}")(n)
    set_config(eng, **cfg)
    eng.set_seed(5)
    dev = torch.device("cuda", 0)
    buf = ds.LeafBuffers(n, n * K, A, oracle.BOARD_SHAPE[game], dev)
    stream = torch.cuda.current_stream().cuda_stream
    buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
    rng = np.random.default_rng(n + npl)
    idx = np.sort(rng.choice(n, sample, replace=False))
    idx[0], idx[-1] = 0, n - 1                                       # first and last tree of the batch
    orc = oracle.OracleMCTS(game, sample)
    set_config(orc, **cfg)
    orc.set_seed(5)
    ev = ev_mod.HashEvaluator(game, mode)
    for mv in range(moves):                                          # moves > 1: the same roots searched again on the kept trees
        ds.playout_device(eng, buf, npl, K, ds.SyntheticEvaluator(game, mode), stream, shards=shards)
        torch.cuda.synchronize()
        c = eng.get_all_counts_array64()
        st = eng.get_all_root_stats()
        # size-independent properties: every simulation below a fresh, non-terminal root passes through exactly one root edge
        assert c.shape == (n, A) and (c >= 0).all()
        assert (c.sum(axis=1) == (mv + 1) * npl - 1).all()
        assert (st[:, 0] == (mv + 1) * npl).all()                    # root_N
        assert np.isfinite(st).all()
        d_p1_p2 = st[:, 3] + st[:, 4] + st[:, 5]                      # root WDL means sum to 1
        assert np.abs(d_p1_p2 - 1.0).max() < 1e-4
        playout(orc, ev, boards[idx], turns[idx], npl, K)
        co = counts(orc, sample, A)
        if not np.array_equal(co, c[idx]):
            bad = np.where((co != c[idx]).any(axis=1))[0]
            raise AssertionError(f"move {mv}: {len(bad)} of {sample} sampled trees differ, first: tree {idx[bad[0]]} "
                                 f"oracle {co[bad[0]]} cuda {c[idx[bad[0]]]}")
        assert orc.get_all_root_stats().tobytes() == np.ascontiguousarray(st[idx]).tobytes()
    return eng


def test_config5_65536_trees_n200_k4_sharded_graph_replay():
    """bench.py's workload (BASELINE config 5 per GPU): 65 536 trees, n=200, K=4, server defaults, symmetry on, 8 shards, the
    native loop replayed from its CUDA graph on the second move."""
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    eng = _run("Connect4", 65536, 200, 4, cfg, "equivariant", 64, None, _c4_roots(65536, 1000), moves=2)
    assert eng.get_lanes() == 1 and eng.get_variant() == 1           # the lean thread-per-tree kernels ran


def test_config5_65536_trees_unsharded_hash_prior():
    """One launch for the whole batch (1024 CTAs, 32-bit chunk indices up to 65 568 * cap * 2), non-equivariant prior."""
    _run("Connect4", 65536, 200, 4, dict(SERVER_DEFAULTS), "hash", 48, 1, _c4_roots(65536, 77))


def test_config3_8192_trees_n800_k8_symmetry_mlh():
    cfg = dict(SERVER_DEFAULTS, c_base=4000.0, use_symmetry=True)
    _run("Connect4", 8192, 800, 8, cfg, "equivariant", 32, None, _c4_roots(8192, 3))


def test_config4_othello_4096_trees_n400_score_utility():
    _run("Othello", 4096, 400, 4, OTH_CFG, "hash", 16, None, _oth_roots(4096, 9))


def test_selfplay_32768_slots_record_invariants():
    """Continuous self-play at scale (tree reuse, arena compaction, slot restarts, temperature sampling): every position of every
    finished game record must be a consistent Connect4 training tuple."""
    import torch
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n, npl, K = 32768, 64, 4
    sp = sp_mod.SelfPlay("Connect4", n, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3),
                         temperature=1.0, temp_decay_moves=12, temp_endgame=0.0, td_steps=5, seed=3, out_capacity=3 * n)
    sp.engine.reserve(4096)
    for _ in range(46):
        sp.ply()
    torch.cuda.synchronize()
    rec = sp.drain()
    m = len(rec)
    assert m >= n, "every slot finishes at least one game in 46 plies"
    assert sp.engine.compactions() > 0
    recs = rec.games
    length, winner, uid = rec.length, rec.winner, rec.uid
    assert int(length.min()) >= 8 and int(length.max()) <= 43 and bool(((winner >= -1) & (winner <= 1)).all())
    assert uid.unique().numel() == m, "game uids are unique"
    rec = rec.sorted_by_uid()
    length, winner, uid = rec.length, rec.winner, rec.uid
    t = rec.to_replay_tensors(5)
    st, prob, mask = t["state"].to(torch.int32), t["prob"], t["valid_mask"]
    P = st.shape[0]
    assert P == int(length.sum())
    own, opp, turn = st[:, 0], st[:, 1], st[:, 2]
    assert bool(((own == 0) | (own == 1)).all()) and bool(((opp == 0) | (opp == 1)).all()) and bool((own * opp == 0).all())
    assert bool((turn.reshape(P, -1).min(1).values == turn.reshape(P, -1).max(1).values).all()) and bool((turn.abs() == 1).all())
    occ = own + opp
    stones = occ.reshape(P, -1).sum(1)
    tsign = turn[:, 0, 0]
    assert bool((tsign == 1 - 2 * (stones % 2)).all()), "player +1 moves on even stone counts"
    assert bool((occ[:, 1:, :] >= occ[:, :-1, :]).all()), "gravity: a stone never floats (row 0 is the top)"
    ste = t["steps_to_end"].reshape(P).to(torch.int32)
    term = ste == 0
    assert int(term.sum()) == m                                                      # one terminal tuple per game
    # non-terminal positions: the legal mask is "column not full", the policy target is a distribution over legal moves
    col_open = occ[:, 0, :] == 0
    assert bool((mask[~term] == col_open[~term]).all())
    assert bool((prob[~term].sum(1) - 1.0).abs().max() < 1e-5) and bool((prob[~term][~mask[~term]] == 0).all())
    assert bool((prob[term] == 0).all())
    # terminal positions: the recorded winner has four in a row (the mover who just played = the opponent plane), else the board is full
    def four(b):
        h = b[:, :, 0:4] * b[:, :, 1:5] * b[:, :, 2:6] * b[:, :, 3:7]
        v = b[:, 0:3] * b[:, 1:4] * b[:, 2:5] * b[:, 3:6]
        d1 = b[:, 0:3, 0:4] * b[:, 1:4, 1:5] * b[:, 2:5, 2:6] * b[:, 3:6, 3:7]
        d2 = b[:, 3:6, 0:4] * b[:, 2:5, 1:5] * b[:, 1:4, 2:6] * b[:, 0:3, 3:7]
        return (h.reshape(len(b), -1).sum(1) + v.reshape(len(b), -1).sum(1) + d1.reshape(len(b), -1).sum(1) + d2.reshape(len(b), -1).sum(1)) > 0
    wz = t["winner"].reshape(P).to(torch.int32)
    last_mover_won = four(opp[term])
    assert bool((~four(own[term])).all()), "the side to move at the end never has four in a row"
    assert bool((last_mover_won == (wz[term] != 0)).all())
    assert bool((wz[term][last_mover_won] == -tsign[term][last_mover_won]).all())
    assert bool((stones[term][~last_mover_won] == 42).all()), "a draw is a full board"
    # within a game, winner_z is constant and steps_to_end counts down to 0
    starts = torch.cumsum(length, 0) - length
    gidx = torch.repeat_interleave(torch.arange(m, device=recs.device), length)
    assert bool((wz == winner[gidx]).all())
    pos_in_game = torch.arange(P, device=recs.device) - starts[gidx]
    assert bool((ste == length[gidx] - 1 - pos_in_game).all())
    assert bool((stones[1:][pos_in_game[1:] > 0] == stones[:-1][pos_in_game[1:] > 0] + 1).all()), "one stone per ply"
