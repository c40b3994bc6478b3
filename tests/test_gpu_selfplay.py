"""On-device self-play driver vs a line-by-line Python restatement of the reference's self-play loop
(src/game.py:65-164 Game.batch_self_play + src/player.py:333-375 get_batch_action) driven on the oracle engine.
With temperature 0 (arg-max) everything is deterministic: every training tuple must match bit for bit."""
import importlib

import numpy as np
import pytest

import oracle
from harness import SERVER_DEFAULTS, counts, playout, random_positions, set_config

pytestmark = pytest.mark.gpu


def reference_style_self_play(engine, evaluator, game, n_games, n_playout, K, td_steps):
    """Restatement of Game.batch_self_play with temp = 0 on any engine with the mcts_cpp surface."""
    A = oracle.ACTION_SIZE[game]
    envs = [oracle.OracleEnv(game) for _ in range(n_games)]
    traj = [dict(states=[], probs=[], players=[], root_wdls=[], masks=[]) for _ in range(n_games)]
    active = list(range(n_games))
    done_data = [None] * n_games
    for i in range(n_games):
        engine.reset_env(i)

    def planes(e):
        b, t = e.board, e.turn
        return np.stack([(b == t), (b == -t), np.full_like(b, t, dtype=np.int8)]).astype(np.int8)

    while active:
        boards = np.stack([e.board for e in envs])
        turns = np.array([e.turn for e in envs], np.int32)
        playout(engine, evaluator, boards, turns, n_playout, K)
        visits = counts(engine, n_games, A)
        root_wdls = engine.get_all_root_stats()[:, 3:6].copy()
        actions, probs = [], []
        for i in range(n_games):                                   # src/player.py:348-371 (temp <= 1e-6 branch)
            v = visits[i]
            p = np.zeros(A, np.float32)
            if not (v > 0).any():
                actions.append(0)
            else:
                p[v > 0] = v[v > 0] / v[v > 0].sum()
                actions.append(int(np.argmax(v)))
            probs.append(p)
        engine.prune_roots(np.array(actions, np.int32))
        nxt = []
        for i in active:                                           # src/game.py:94-160
            e, t = envs[i], traj[i]
            t["states"].append(planes(e)); t["probs"].append(probs[i]); t["root_wdls"].append(root_wdls[i])
            mask = np.zeros(A, bool); mask[e.valid_moves()] = True
            t["masks"].append(mask); t["players"].append(e.turn)
            e.step(actions[i])
            if e.done():
                T = len(t["players"])
                winner = e.winner()
                ste = np.arange(T, 0, -1, dtype=np.int32)
                if game == "Othello":
                    diff = int(np.sum(e.board == 1) - np.sum(e.board == -1))
                    aux, term_aux = diff * np.asarray(t["players"], np.int32), diff * e.turn
                else:
                    aux, term_aux = ste, 0
                fut = [t["root_wdls"][k + td_steps] if k + td_steps < T else np.zeros(3, np.float32) for k in range(T)]
                done_data[i] = dict(winner=winner, length=T + 1, state=np.stack(t["states"] + [planes(e)]),
                                    prob=np.stack(t["probs"] + [np.zeros(A, np.float32)]),
                                    root_wdl=np.stack(t["root_wdls"] + [np.zeros(3, np.float32)]),
                                    future_root_wdl=np.stack(fut + [np.zeros(3, np.float32)]),
                                    winner_z=np.full(T + 1, winner, np.int32), steps_to_end=np.append(ste, 0),
                                    aux=np.append(aux, term_aux).astype(np.int32),
                                    valid_mask=np.stack(t["masks"] + [np.ones(A, bool)]))
                engine.reset_env(i)
            else:
                nxt.append(i)
        active = nxt
    return done_data


@pytest.mark.parametrize("game,n,npl,K,cfg", [
    ("Connect4", 48, 40, 4, dict(SERVER_DEFAULTS, use_symmetry=True)),
    ("Connect4", 32, 25, 1, dict(SERVER_DEFAULTS)),
    ("Othello", 12, 24, 4, dict(c_init=1.4, c_base=500.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=True,
                                score_utility_factor=0.15, score_scale=8.0)),
])
def test_selfplay_records_equal_reference_style_loop(game, n, npl, K, cfg):
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    td = 3
    orc = oracle.OracleMCTS(game, n)
    set_config(orc, **cfg)
    orc.set_seed(11)
    want = reference_style_self_play(orc, ev_mod.HashEvaluator(game, "hash"), game, n, npl, K, td)
    sp = sp_mod.SelfPlay(game, n, npl, K, ds.SyntheticEvaluator(game, "hash"), search_cfg=cfg, temperature=0.0, temp_decay_moves=0,
                         temp_endgame=0.0, td_steps=td, seed=11, out_capacity=8 * n)
    recs = sp.run(target_games=8 * n, max_plies=max(w["length"] for w in want) + 1)
    got = {g["uid"]: g for g in recs.unpack(td) if g["uid"] < n}
    assert sorted(got) == list(range(n)), "every first-generation game must have been flushed"
    for i in range(n):
        g, w = got[i], want[i]
        assert g["winner"] == w["winner"] and g["length"] == w["length"], f"game {i}"
        for k in ("state", "prob", "root_wdl", "future_root_wdl", "winner_z", "steps_to_end", "aux", "valid_mask"):
            assert np.array_equal(np.asarray(g[k]), w[k]), f"game {i} field {k}"
    w0, tup = got[0]["tuples"]
    assert w0 == want[0]["winner"] and len(tup) == want[0]["length"] and len(tup[0]) == 8 and tup[0][0].dtype == np.int8


@pytest.mark.parametrize("name", ["py_c4_selfplay_k4_sym", "py_c4_selfplay_k1_td0", "py_oth_selfplay_k4"])
def test_selfplay_driver_reproduces_reference_batch_self_play(name):
    """Pinned to the reference: tests/golden/py_*_selfplay_* hold what the UNMODIFIED Game.batch_self_play + get_batch_action
    returned on the compiled reference engine (temperature 1 on the opening plies, sampled from numpy's generator).  The on-device
    driver replays the recorded moves (opening script; the sampling RNG itself is unpinned by nature) and must produce the same
    compact records bit for bit - policy targets from its own visit counts, root WDL from its own root statistics, winner, length,
    positions - hence, through the expansion kernel, every training tuple of the fixture."""
    import json
    import os
    import torch
    from fixture_records import GOLD, fixture_records, forced_actions, load_fixture
    from test_record_formats import _assert_tensors_match_fixture
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    case = json.load(open(os.path.join(GOLD, "py_cases.json")))["selfplay"][name]
    z = load_fixture(name)
    game, n, td = case["game"], case["n_games"], case["td_steps"]
    cfg = dict(case["cfg"], dirichlet_alpha=0.0, use_symmetry=case["use_symmetry"])
    cfg.setdefault("mlh_slope", 0.0)
    sp = sp_mod.SelfPlay(game, n, case["n_playout"], case["K"], ds.SyntheticEvaluator(game, case["mode"]), search_cfg=cfg,
                         temperature=case["temperature"], temp_decay_moves=case["temp_decay_moves"], temp_endgame=0.0, td_steps=td,
                         seed=case["seed"], out_capacity=8 * n, forced_actions=forced_actions(z, game), forced_uid0=0)
    recs = sp.run(target_games=8 * n, max_plies=int(z["length"].max()))
    first = recs.uid < n
    mine = sp_mod.Records.cat([sp_mod.Records(game, recs.games[i:i + 1], recs.pos) for i in torch.nonzero(first).flatten().tolist()]).sorted_by_uid()
    want = fixture_records(z, game)
    assert mine.uid.tolist() == list(range(n)) and mine.length.tolist() == want.length.tolist() and mine.winner.tolist() == want.winner.tolist()
    assert torch.equal(mine.games.cpu(), want.games), "game headers differ from the reference's games"
    assert torch.equal(mine.pos.cpu(), want.pos), "position records differ (bitboards / side to move / root WDL / policy target)"
    _assert_tensors_match_fixture(mine.to_replay_tensors(td), z, td)


def test_self_play_records_do_not_depend_on_the_sharding():
    """game i must give the same record whichever rank / slot range owns it.  One engine with G slots against
    two engines with G/2 slots each (`set_env_base`, `uid_base`, `uid_stride` = G, as rank 0 / rank 1 of a 2-GPU run would be set up),
    everything RNG-dependent switched on (Dirichlet noise, random leaf symmetries, temperature sampling): identical records per uid."""
    import torch
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    G, npl, K, plies = 256, 40, 4, 30
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3)

    def play(n, base):
        sp = sp_mod.SelfPlay("Connect4", n, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=cfg, temperature=1.0, temp_decay_moves=8,
                             td_steps=4, seed=21, uid_base=base, uid_stride=G, out_capacity=6 * n)
        for _ in range(plies):
            sp.ply()
        return sp.drain()

    whole = play(G, 0).sorted_by_uid()
    halves = sp_mod.Records.cat([play(G // 2, 0), play(G // 2, G // 2)]).sorted_by_uid()
    assert len(whole) > G and whole.uid.unique().numel() == len(whole)
    assert torch.equal(whole.games, halves.games) and torch.equal(whole.pos, halves.pos)
    t = whole.to_replay_tensors(4)
    assert float(t["prob"].sum()) > 0


def test_temperature_sampling_follows_visit_distribution():
    """RNG-dependent (parity unpinned): with temp = 1 the played move is distributed like visits/sum."""
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n = 8192
    sp = sp_mod.SelfPlay("Connect4", n, 64, 4, ds.SyntheticEvaluator("Connect4", "constant"),
                         search_cfg=dict(SERVER_DEFAULTS, use_symmetry=False), temperature=1.0, temp_decay_moves=100, td_steps=0, seed=5)
    sp.ply()
    import torch
    torch.cuda.synchronize()
    probs = sp.st_pos[:, 0, 32:60].contiguous().view(torch.float32).cpu().numpy()      # identical trees: same visit distribution in every slot
    acts = sp.actions.cpu().numpy()
    assert np.allclose(probs, probs[0])
    freq = np.bincount(acts, minlength=7) / n
    assert np.abs(freq - probs[0]).max() < 0.02


class _RowDeterministicNet:
    """A 'network' whose output for a row depends only on that row (exact integer arithmetic), so evaluating a subset of
    the batch gives bit-identical rows - lets the cache be tested for exact equality."""
    def predict_device(self, planes, mask):
        import torch
        B = planes.shape[0]
        w = torch.arange(1, 43, device=planes.device, dtype=torch.int64)
        key = (planes[:, 0].reshape(B, 42).long() * w).sum(1) * 7919 + (planes[:, 1].reshape(B, 42).long() * w).sum(1) * 104729 \
            + (planes[:, 2, 0, 0].long() + 2) * 1299709
        a = torch.arange(7, device=planes.device, dtype=torch.int64)[None, :]
        probs = (((key[:, None] * 31 + a * 977) % 4096) + 1).float() / 4096.0
        w3 = torch.stack([((key * 13) % 251 + 1), ((key * 17) % 241 + 1), ((key * 19) % 239 + 1)], 1).float()
        return probs, w3 / w3.sum(1, keepdim=True), ((key % 40) + 1).float()


def test_device_evaluation_cache_is_transparent():
    """Cached evaluations are exactly what the network returns, so visit counts with and without the cache are identical,
    while most leaves of a self-play batch (same openings in every slot) never reach the network."""
    import torch
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    n, npl, K = 256, 60, 4
    boards = np.zeros((n, 6, 7), np.int8)
    turns = np.ones(n, np.int32)
    kw = dict(game_name="Connect4", noise_epsilon=0.0, fpu_reduction=0.2, use_symmetry=True, mlh_slope=0.1)
    plain = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, npl, cache_size=0, **kw)
    cached = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, npl, cache_size=10000, **kw)
    for e in (plain, cached):
        e.seed(3)
    net = _RowDeterministicNet()
    net_rows = 0
    for mv in range(3):
        plain.batch_playout(net, boards, turns, vl_batch=K)
        cached.batch_playout(net, boards, turns, vl_batch=K)
        net_rows += cached._last_evaluator.net_rows
        a, b = plain.get_visits_count(), cached.get_visits_count()
        assert np.array_equal(a, b)
        acts = a.argmax(1).astype(np.int32)
        for i in range(n):                                   # play the move on the host boards
            r = int(np.max(np.where(boards[i, :, acts[i]] == 0)[0]))
            boards[i, r, acts[i]] = turns[i]
        turns = -turns
        plain.prune_roots(acts)
        cached.prune_roots(acts)
    st = cached._dev_cache.stats()
    # in-batch duplicates share one evaluation (the reference evaluates each): the identical openings of the first batches
    assert st["hits"] + st["dups"] > 0.5 * st["lookups"] and st["inserts"] > 0 and st["dups"] > 0, st
    assert net_rows == st["lookups"] - st["hits"] - st["dups"], (net_rows, st)
    cached.refresh_cache(net)                                # weight reload: the device cache is dropped
    cached.batch_playout(net, boards, turns, vl_batch=K)
    st2 = cached._dev_cache.stats()
    assert st2["inserts"] > st["inserts"]


@pytest.mark.parametrize("log2cap", [4, 12])
def test_cache_dedup_on_leaf_records(log2cap):
    """az_evalcache_lookup_dedup_dev / insert / resolve on hand-made leaf records: repeated positions, terminal rows and (in the
    16-entry table) many different positions per entry.  Every non-terminal row ends up with the outputs of its own position,
    a distinct missing position is evaluated once unless it shares its table entry with another one, and a second batch hits whatever was stored."""
    import ctypes as C
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    L = importlib.import_module("alphazero-al_b200._lib").lib()
    rng = np.random.default_rng(11)
    n, distinct = 4096, 300
    pos = rng.integers(1, 1 << 40, size=(distinct, 2), dtype=np.uint64)
    pick = rng.integers(0, distinct, size=n)
    rec = np.zeros(n, dtype=[("bb0", "<u8"), ("bb1", "<u8"), ("turn", "i1"), ("flags", "u1"), ("sym", "u1"), ("passes", "u1"), ("r", "<i4", 3)])
    rec["bb0"], rec["bb1"] = pos[pick, 0], pos[pick, 1]
    rec["turn"] = np.where(pick % 2 == 0, 1, -1)
    term = rng.random(n) < 0.1
    rec["flags"] = term.astype(np.uint8)
    assert rec.dtype.itemsize == 32
    dev = torch.device("cuda:0")
    leaves = torch.from_numpy(rec.view(np.uint8).reshape(n, 32)).to(dev)

    def net_rows(idx):                                       # the "network": a pure function of the position
        k = torch.as_tensor(pick, device=dev)[idx].float()
        return ((k[:, None] + torch.arange(7, device=dev)[None, :]) / 512.0).contiguous(), \
               torch.stack([k, k + 0.25, k + 0.5], 1).contiguous(), (k * 3 + 1).contiguous()

    cache = ds.EvalCache("Connect4", (1 << log2cap) // 2)
    assert cache.stats()["capacity"] == 1 << log2cap
    for batch in range(2):
        probs = torch.full((n, 7), -1.0, device=dev)
        wdl = torch.full((n, 3), -1.0, device=dev)
        aux = torch.full((n,), -1.0, device=dev)
        miss_idx = torch.empty(n, dtype=torch.int32, device=dev)
        miss_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        dup_of = torch.full((n,), -7, dtype=torch.int32, device=dev)
        assert L.az_evalcache_lookup_dedup_dev(cache._c, n, leaves.data_ptr(), probs.data_ptr(), wdl.data_ptr(), aux.data_ptr(),
                                               miss_idx.data_ptr(), miss_cnt.data_ptr(), dup_of.data_ptr(), None) == 0
        m = int(miss_cnt.item())
        idx = miss_idx[:m].long()
        missed = pick[idx.cpu().numpy()]
        if log2cap == 12 and batch == 0:                     # positions sharing a table entry with another one are not de-duplicated
            assert m <= 2 * len(set(missed.tolist())) and m < 0.2 * n
        assert not term[idx.cpu().numpy()].any()
        if batch == 0:
            assert set(missed.tolist()) == set(pick[~term].tolist())
        pm, wm, am = net_rows(idx)
        assert L.az_evalcache_insert_dev(cache._c, m, leaves.data_ptr(), miss_idx.data_ptr(), pm.data_ptr(), wm.data_ptr(), am.data_ptr(),
                                         probs.data_ptr(), wdl.data_ptr(), aux.data_ptr(), None) == 0
        assert L.az_evalcache_resolve_dups_dev(cache._c, n, dup_of.data_ptr(), probs.data_ptr(), wdl.data_ptr(), aux.data_ptr(), None) == 0
        torch.cuda.synchronize()
        live = torch.as_tensor(~term, device=dev)
        wp, ww, wa = net_rows(torch.arange(n, device=dev))
        assert torch.equal(probs[live], wp[live]) and torch.equal(wdl[live], ww[live]) and torch.equal(aux[live], wa[live])
        assert bool((probs[~live] == -1).all()) and bool((dup_of[~live] == -1).all())
        d = dup_of.cpu().numpy()
        own = d[d >= 0]
        assert (pick[own] == pick[d >= 0]).all() and (d[own] == -1).all()
    st = cache.stats()
    assert st["lookups"] == 2 * int((~term).sum()) and st["dups"] > 0
    if log2cap == 12:
        assert st["hits"] >= int((~term).sum()) * 0.9        # second batch: nearly everything was stored by the first
    assert st["inserts"] <= st["lookups"] - st["hits"] - st["dups"]


def test_net_evaluator_graph_replay_equals_eager():
    """Small batches replay the network from a captured CUDA graph (NetEvaluator.graph_rows); the search must not notice."""
    import torch
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    nets = importlib.import_module("alphazero-al_b200.nets")
    n, npl, K = 96, 41, 4
    boards, turns = random_positions("Connect4", n, 12, seed=21)
    kw = dict(game_name="Connect4", noise_epsilon=0.0, fpu_reduction=0.2, use_symmetry=True, mlh_slope=0.1)
    for net in (_RowDeterministicNet(), None):
        if net is None:
            torch.manual_seed(4)
            net = nets.C4Net(device="cuda:0")
            for head in (net.p_out, net.v_wdl, net.v_aux):                     # non-trivial outputs
                torch.nn.init.normal_(head.weight, std=0.5)
        res = []
        for graph_rows in (0, 8192):
            eng = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, npl, **kw)
            eng.seed(9)
            for mv in range(2):
                eng.batch_playout(net, boards, turns, vl_batch=K)
                ev = eng._last_evaluator
                if mv == 0:
                    ev.graph_rows = graph_rows                                  # takes effect from the second move on
                    if graph_rows == 0:
                        ev._graphs.clear()
                cnt = eng.get_visits_count()
                eng.prune_roots(np.full(n, -1, np.int32))
            res.append((cnt, ev.graph_replays))
        (c0, r0), (c1, r1) = res
        assert r1 > r0
        if isinstance(net, _RowDeterministicNet):
            assert np.array_equal(c0, c1)
        else:                                                                   # same kernels, same inputs: equal in practice; L1 bound as in north_star
            p0, p1 = c0 / c0.sum(1, keepdims=True), c1 / c1.sum(1, keepdims=True)
            assert np.abs(p0 - p1).sum(1).max() <= 1e-3


def test_reference_style_network_takes_the_device_path():
    """A torch module with the reference networks' forward contract (forward(state, action_mask) -> log-policy, log-WDL,
    normalised aux; `aux_target_offset`; numpy `predict`) and weights on the GPU is wrapped by batch_playout itself
    (ReferenceNetAdapter) and gives the trees its own predict() gives through the host path."""
    import torch
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    nets = importlib.import_module("alphazero-al_b200.nets")

    class RefLike(torch.nn.Module):
        aux_target_offset = 42

        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, action_mask=None):
            return self.inner(x, action_mask)

        def predict(self, state, action_mask=None):
            return self.inner.predict(state, action_mask)

    torch.manual_seed(6)
    inner = nets.C4Net(device="cuda:0")
    for head in (inner.p_out, inner.v_wdl, inner.v_aux):
        torch.nn.init.normal_(head.weight, std=0.5)
    ref_like = RefLike(inner).eval()
    assert ds.ReferenceNetAdapter.accepts(ref_like) and not ds.ReferenceNetAdapter.accepts(inner)
    n, npl, K = 64, 41, 4
    boards, turns = random_positions("Connect4", n, 12, seed=8)
    kw = dict(game_name="Connect4", noise_epsilon=0.0, fpu_reduction=0.2, use_symmetry=False, mlh_slope=0.1)
    a, b = bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, npl, **kw), bm.BatchedMCTS(n, 1.4, 1000.0, 0.0, npl, **kw)
    a.batch_playout(ref_like, boards, turns, vl_batch=K)
    assert isinstance(a._last_evaluator.net, ds.ReferenceNetAdapter)
    b.batch_playout(inner, boards, turns, vl_batch=K)
    ca, cb = a.get_visits_count(), b.get_visits_count()
    pa, pb = ca / ca.sum(1, keepdims=True), cb / cb.sum(1, keepdims=True)
    assert np.abs(pa - pb).sum(1).max() <= 1e-3


def test_output_ring_overflow_is_reported_not_silent():
    """games finished after the ring filled up used to vanish.  Now every finished game is either in the ring or counted as
    dropped, and drain() raises instead of returning a truncated set; draining in time loses nothing."""
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    n = 256
    kw = dict(search_cfg=dict(SERVER_DEFAULTS), temperature=1.0, temp_decay_moves=50, td_steps=0, seed=2)
    sp = sp_mod.SelfPlay("Connect4", n, 16, 4, ds.SyntheticEvaluator("Connect4", "constant"), out_capacity=64, **kw)
    for _ in range(43):                                       # every slot finishes at least one game: 256 > 64
        sp.ply()
    with pytest.raises(RuntimeError, match="overflowed"):
        sp.drain()
    sp2 = sp_mod.SelfPlay("Connect4", n, 16, 4, ds.SyntheticEvaluator("Connect4", "constant"), out_capacity=n, **kw)   # >= one ply's worth
    total, uids = 0, set()
    for _ in range(43):
        sp2.ply()
        if sp2.finished() > 16:
            rec = sp2.drain()
            total += len(rec)
            uids.update(rec.uid.tolist())
    rec = sp2.drain()
    total += len(rec)
    uids.update(rec.uid.tolist())
    assert total >= n and len(uids) == total and set(range(n)) <= uids
    again = sp2.run(target_games=10)                          # run() after a drain continues with new games (no stale records)
    assert len(again) >= 10 and not (set(again.uid.tolist()) & uids)
