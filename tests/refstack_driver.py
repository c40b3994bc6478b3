"""Runs the reference's UNMODIFIED Python layer (whatever ``src`` package is first on PYTHONPATH: see oracle/refstack.py) and dumps
what it produced.  TEST INFRASTRUCTURE: used by tests/golden/make_golden_py.py (engine = compiled reference -> committed fixtures),
by the ``-m gpu`` drop-in tests (engine = this repository's modules -> compared with those fixtures) and by bench.py's
reference-actor leg.

    python refstack_driver.py <task> <out path> '<json params>'

tasks
  selfplay  Game.batch_self_play (src/game.py:65-164) with AlphaZeroPlayer.get_batch_action (src/player.py:333-375) -> out.npz
            (flattened training tuples); with "formats": also out.pt (ReplayBuffer.store + save, src/ReplayBuffer.py:25-39,89-121,
            fed like server.py:295-304) and out.pkl (the upload payload of client.py:367-373)
  playout   src.MCTS_cpp.BatchedMCTS.batch_playout over several moves with the LRU cache on -> counts / root stats per move, the
            cache's key order and values at the end
  load_pt   the reference's ReplayBuffer.load on a .pt file -> what it holds afterwards
  cnn       the reference's own CNN (bf16 autocast on CUDA) in the loop: (a) its predict() drives the compiled reference engine through
            the reference-style host loop while every leaf batch and evaluator output is recorded, (b) the CUDA engine is fed the
            recorded outputs (same evaluator outputs to both -> leaves and visit counts must be identical), (c) the
            device-resident path (ReferenceNetAdapter, no host copy) is compared with (a) in L1, (d) the adapter's outputs are
            compared with predict() on one batch -> out.json
  actor     wall-clock of one batch_self_play call with the reference's own CNN (config 1 of BASELINE.json) -> out.json
"""
import json
import pickle
import sys
import time

import numpy as np


class HashPredict:
    """The deterministic hash evaluator behind the reference's ``predict(state, action_mask=None)`` contract
    (src/environments/Connect4/Network.py:267-288): state f32[B,3,R,C] relative planes -> (probs[B,A], wdl_rel[B,3], aux[B,1])."""

    def __init__(self, game, mode):
        import importlib
        self.ev = importlib.import_module("alphazero-al_b200.evaluators").HashEvaluator(game, mode)
        self.n_actions = self.ev.A
        self.calls = self.rows = 0

    def predict(self, state, action_mask=None):
        state = np.asarray(state)
        turn = state[:, 2, 0, 0].astype(np.int32)
        board = ((state[:, 0] - state[:, 1]) * turn[:, None, None]).astype(np.int8)
        probs, wdl, aux = self.ev.raw(board, turn)
        self.calls += 1
        self.rows += len(board)
        return probs, wdl, aux.reshape(-1, 1)

    def eval(self):
        return self

    def train(self):
        return self


def flatten_games(completed):
    """[(winner, play_data)] -> dict of arrays (positions of all games concatenated, `length[i]` rows per game)."""
    width = len(completed[0][1][0])
    names = ["state", "prob", "winner_z", "steps_to_end", "aux", "root_wdl", "valid_mask", "future_root_wdl"][:width]
    out = {"winner": np.array([w for w, _ in completed], np.int32), "length": np.array([len(pd) for _, pd in completed], np.int32),
           "tuple_width": np.array(width, np.int32)}
    for j, nm in enumerate(names):
        out[nm] = np.stack([np.asarray(row[j]) for _, pd in completed for row in pd])
    # the move played at every position (-1 at the terminal tuple), recovered from consecutive states: the one cell that was empty
    # and is occupied afterwards (Othello: none -> pass = 64)
    acts = []
    for _, pd in completed:
        for t in range(len(pd)):
            if t + 1 == len(pd):
                acts.append(-1)
                continue
            occ0 = (pd[t][0][0] + pd[t][0][1]).reshape(-1)
            occ1 = (pd[t + 1][0][0] + pd[t + 1][0][1]).reshape(-1)
            new = np.where((occ0 == 0) & (occ1 != 0))[0]
            cols = pd[t][0].shape[2]
            acts.append(int(new[0]) % cols if cols == 7 else (int(new[0]) if len(new) else 64))
    out["action"] = np.array(acts, np.int32)
    return out


def type_signature(play_data):
    """Python / numpy types of the first and of the terminal tuple of a game (what pickles over the wire)."""
    def sig(row):
        return [f"{type(x).__module__}.{type(x).__name__}:{getattr(x, 'dtype', '')}:{getattr(x, 'shape', '')}" for x in row]
    return sig(play_data[0]), sig(play_data[-1])


def make_player(p, pv):
    from src.player import AlphaZeroPlayer
    cfg = p.get("cfg", {})
    return AlphaZeroPlayer(pv, n_envs=p["n_games"], c_init=cfg.get("c_init", 1.4), c_base=cfg.get("c_base", 1000), n_playout=p["n_playout"],
                           alpha=p.get("alpha", 0.0), is_selfplay=1, cache_size=p.get("cache_size", 0), noise_epsilon=cfg.get("noise_epsilon", 0.25),
                           fpu_reduction=cfg.get("fpu_reduction", 0.2), use_symmetry=p.get("use_symmetry", False), game_name=p["game"],
                           mlh_slope=cfg.get("mlh_slope", 0.0), mlh_cap=cfg.get("mlh_cap", 0.2),
                           score_utility_factor=cfg.get("score_utility_factor", 0.0), score_scale=cfg.get("score_scale", 8.0),
                           value_decay=cfg.get("value_decay", 1.0), vl_batch=p["K"])


def task_selfplay(out, p):
    from src.environments import load
    from src.game import Game
    mod = load(p["game"])
    pv = HashPredict(p["game"], p.get("mode", "hash"))
    player = make_player(p, pv)
    player.mcts.seed(p.get("seed", 0))
    np.random.seed(p.get("seed", 0))
    game = Game(mod.Env())
    done = game.batch_self_play(player, p["n_games"], p.get("temperature", 0.0), p.get("temp_decay_moves", 0), p.get("temp_endgame", 0.0),
                                td_steps=p["td_steps"])
    flat = flatten_games(done)
    first, last = type_signature(done[0][1])
    flat["sig_first"], flat["sig_last"] = np.array(first), np.array(last)
    flat["env_module"] = np.array(type(game.env).__module__)
    flat["engine_module"] = np.array(type(player.mcts.mcts).__module__)
    np.savez_compressed(out, **flat)
    if p.get("formats"):
        from src.ReplayBuffer import ReplayBuffer
        R, Cc = done[0][1][0][0].shape[1:]
        total = int(flat["length"].sum())
        buf = ReplayBuffer(3, total, pv.n_actions, R, Cc)
        for _, play_data in done:                              # server.py:295-304 inbox_worker
            for data in play_data:
                buf.store(*data)
        buf.save(out[:-4] + ".pt")
        payload = pickle.dumps({"__az__": True, "data": [pd for _, pd in done]}, protocol=pickle.HIGHEST_PROTOCOL)      # client.py:367-368
        with open(out[:-4] + ".pkl", "wb") as f:
            f.write(payload)
    print("selfplay", p["game"], "games", len(done), "positions", int(flat["length"].sum()), "predict calls", pv.calls)


def task_playout(out, p):
    from harness import random_positions
    if p.get("wrapper", "src") == "mirror":                    # this repository's mirror of the wrapper instead of the reference's file
        import importlib
        BatchedMCTS = importlib.import_module("alphazero-al_b200.batched_mcts").BatchedMCTS
    else:
        from src.MCTS_cpp import BatchedMCTS
    cfg = p.get("cfg", {})
    n = p["n"]
    pv = HashPredict(p["game"], p.get("mode", "hash"))
    eng = BatchedMCTS(n, cfg.get("c_init", 1.4), cfg.get("c_base", 1000), 0.0, p["n_playout"], game_name=p["game"], cache_size=p["cache_size"],
                      noise_epsilon=0.0, fpu_reduction=cfg.get("fpu_reduction", 0.2), use_symmetry=p.get("use_symmetry", False),
                      mlh_slope=cfg.get("mlh_slope", 0.0), mlh_cap=cfg.get("mlh_cap", 0.2),
                      score_utility_factor=cfg.get("score_utility_factor", 0.0), score_scale=cfg.get("score_scale", 8.0),
                      value_decay=cfg.get("value_decay", 1.0))
    eng.seed(p.get("seed", 0))
    import oracle
    boards, turns = random_positions(p["game"], n, p.get("max_plies", 10), p.get("pos_seed", 1))
    envs = [oracle.OracleEnv(p["game"]) for _ in range(n)]
    for i, e in enumerate(envs):
        e.import_board(boards[i], turns[i])
    res = {"boards": boards, "turns": turns}
    for mv in range(p["moves"]):
        b = np.stack([e.board for e in envs]).astype(np.float32)          # Env.board is float32 in the reference (env_common.h)
        t = np.array([e.turn for e in envs], np.int32)
        eng.batch_playout(pv, b, t, vl_batch=p["K"])
        c = eng.get_visits_count()
        st = eng.get_root_stats()
        res[f"counts{mv}"] = np.asarray(c)
        for k, v in st.items():
            res[f"stats{mv}_{k}"] = np.asarray(v)
        res[f"probs{mv}"] = np.asarray(eng.get_mcts_probs())
        acts = np.zeros(n, np.int32)
        for i, e in enumerate(envs):
            if e.done() or c[i].sum() == 0:
                e.reset()
                eng.reset_env(i)
                acts[i] = -1
            else:
                acts[i] = int(np.argmax(c[i]))
                e.step(acts[i])
        eng.prune_roots(acts)
        res[f"actions{mv}"] = acts
        res[f"predict_rows{mv}"] = np.array(pv.rows, np.int64)
    od = eng.cache._od
    keys = list(od.keys())
    res["cache_keys"] = np.frombuffer(b"".join(keys), np.uint8).reshape(len(keys), -1) if keys else np.zeros((0, 1), np.uint8)
    res["cache_probs"] = np.stack([od[k]["value"][0] for k in keys])
    res["cache_wdl"] = np.stack([od[k]["value"][1] for k in keys])
    res["cache_ml"] = np.array([od[k]["value"][2] for k in keys], np.float64)
    res["counts_dtype"] = np.array(str(np.asarray(c).dtype))
    np.savez_compressed(out, **res)
    print("playout", p["game"], "moves", p["moves"], "cache entries", len(keys), "predict rows", pv.rows)


def task_load_pt(out, p):
    """The reference's ReplayBuffer.load on a file (src/ReplayBuffer.py:40-62)."""
    from src.ReplayBuffer import ReplayBuffer
    buf = ReplayBuffer(3, p["rows"], p["A"], p["R"], p["C"])
    buf.load(p["path"])
    np.savez_compressed(out, ptr=np.array(buf._ptr), len=np.array(len(buf)), state=buf.state.numpy(), prob=buf.prob.numpy(),
                        winner=buf.winner.numpy(), steps_to_end=buf.steps_to_end.numpy(), aux_target=buf.aux_target.numpy(),
                        root_wdl=buf.root_wdl.numpy(), valid_mask=buf.valid_mask.numpy(), future_root_wdl=buf.future_root_wdl.numpy())
    print("load_pt", len(buf))


def task_cnn(out, p):
    import importlib
    import torch
    import oracle
    from harness import SERVER_DEFAULTS, counts, playout, random_positions, set_config
    from src.environments import load
    from src import mcts_cpp as ref_cpp                          # overlay "reference": the compiled reference engine
    ours_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
    bm = importlib.import_module("alphazero-al_b200.batched_mcts")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    game, n, npl, K = p["game"], p["n"], p["n_playout"], p["K"]
    A = oracle.ACTION_SIZE[game]
    mod = load(game)
    torch.manual_seed(p.get("seed", 0))
    net = mod.CNN(lr=0.0, device="cuda")
    net.eval()
    with torch.no_grad():                                        # random-init heads are nearly uniform: give them some signal
        for q in net.parameters():
            q.add_(p.get("perturb", 0.05) * torch.randn_like(q))
    boards, turns = random_positions(game, n, p.get("max_plies", 14), p.get("pos_seed", 41))
    cfg = dict(SERVER_DEFAULTS, use_symmetry=False)
    if game == "Othello":
        cfg = dict(c_init=1.4, c_base=500.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False, score_utility_factor=0.15, score_scale=8.0)
    log = []

    def host_eval(lb, lt, it, td, tp1, tp2, vm):                 # src/MCTS_cpp.py:275-297 with the network's own predict()
        t = it.astype(bool)
        probs = np.zeros((lb.shape[0], A), np.float32)
        d, p1w, p2w, ml = td.copy(), tp1.copy(), tp2.copy(), np.zeros(lb.shape[0], np.float32)
        if (~t).any():
            planes = bm._default_convert_board(lb[~t], lt[~t])
            mask = vm[~t].astype(bool, copy=False)
            pr, w, a = net.predict(planes, mask)
            probs[~t] = pr
            d[~t] = w[:, 0]
            p1w[~t] = np.where(lt[~t] == 1, w[:, 1], w[:, 2])
            p2w[~t] = np.where(lt[~t] == 1, w[:, 2], w[:, 1])
            ml[~t] = a.reshape(-1)
        log.append((lb.copy(), lt.copy(), it.copy(), probs, d, p1w, p2w, ml, vm.copy()))
        return probs, d, p1w, p2w, ml

    host_eval.wants_mask = True
    ref = getattr(ref_cpp, f"BatchedMCTS_{game}")(n)
    set_config(ref, **cfg)
    playout(ref, host_eval, boards, turns, npl, K)
    theirs = counts(ref, n, A)
    # (b) oracle-fed: same evaluator outputs to the CUDA engine; its leaves must be the recorded ones at every iteration
    it_no = [0]

    def replay_eval(lb, lt, it, td, tp1, tp2):
        rec = log[it_no[0]]
        it_no[0] += 1
        assert np.array_equal(lb, rec[0]) and np.array_equal(lt, rec[1]) and np.array_equal(it, rec[2]), f"leaves differ at iteration {it_no[0] - 1}"
        return rec[3:8]

    mine = getattr(ours_cpp, f"BatchedMCTS_{game}")(n)
    set_config(mine, **cfg)
    playout(mine, replay_eval, boards, turns, npl, K)
    fed = counts(mine, n, A)
    fed_equal = bool(np.array_equal(fed, theirs)) and mine.get_all_root_stats().tobytes() == ref.get_all_root_stats().tobytes()

    def l1(a, b):
        a, b = a.astype(np.float64), b.astype(np.float64)
        return np.abs(a / a.sum(1, keepdims=True) - b / b.sum(1, keepdims=True)).sum(1)

    # (c) the device-resident path: batch_playout wraps the CUDA reference network by itself (ReferenceNetAdapter)
    kw = dict(game_name=game, noise_epsilon=0.25, fpu_reduction=cfg["fpu_reduction"], use_symmetry=False, mlh_slope=cfg.get("mlh_slope", 0.0),
              mlh_cap=cfg.get("mlh_cap", 0.2), score_utility_factor=cfg.get("score_utility_factor", 0.0), score_scale=cfg.get("score_scale", 8.0))
    dev_eng = bm.BatchedMCTS(n, cfg["c_init"], cfg["c_base"], 0.0, npl, **kw)
    dev_eng.batch_playout(net, boards, turns, vl_batch=K)
    adapter_used = isinstance(dev_eng._last_evaluator.net, ds.ReferenceNetAdapter)
    l1_dev = l1(dev_eng.get_visits_count(), theirs)
    # (d) adapter vs predict() on the first recorded non-terminal leaf batch
    lb, lt, it = log[1][0], log[1][1], log[1][2].astype(bool)
    planes = bm._default_convert_board(lb[~it], lt[~it])
    masks = log[1][8][~it].astype(bool)
    net.score_scale = cfg.get("score_scale", 8.0)
    pr, w, a = net.predict(planes, masks)
    ad = ds.ReferenceNetAdapter(net, game)
    pd_, wd_, ad_ = ad.predict_device(torch.from_numpy(planes.astype(np.float32)).cuda(), torch.from_numpy(masks.astype(np.uint8)).cuda())
    adapter_equal = bool(np.array_equal(pd_.cpu().numpy(), pr) and np.array_equal(wd_.cpu().numpy(), w) and np.array_equal(ad_.cpu().numpy(), a.reshape(-1)))
    adapter_maxdiff = float(max(np.abs(pd_.cpu().numpy() - pr).max(), np.abs(wd_.cpu().numpy() - w).max(), np.abs(ad_.cpu().numpy() - a.reshape(-1)).max()))
    res = dict(game=game, iterations=len(log), fed_counts_and_stats_identical=fed_equal, fed_l1_max=float(l1(fed, theirs).max()),
               device_path_l1_max=float(l1_dev.max()), device_path_l1_mean=float(l1_dev.mean()), device_path_identical_trees=float(np.mean(l1_dev == 0)),
               adapter_used=adapter_used, adapter_equals_predict=adapter_equal, adapter_maxdiff=adapter_maxdiff, net=type(net).__module__,
               policy_spread=float(pr.max() - pr.min()))
    with open(out, "w") as f:
        json.dump(res, f)
    print("cnn", json.dumps(res))


def task_actor(out, p):
    import torch
    from src.environments import load
    from src.game import Game
    mod = load(p["game"])
    torch.manual_seed(0)
    np.random.seed(0)
    dev = p.get("device", "cuda" if torch.cuda.is_available() else "cpu")
    net = mod.CNN(lr=0.0, device=dev) if "device" in mod.CNN.__init__.__code__.co_varnames else mod.CNN(lr=0.0)
    net.eval()
    q = dict(p, cfg=dict(c_init=1.4, c_base=1000, fpu_reduction=0.2, mlh_slope=0.1, mlh_cap=0.2), alpha=0.3, use_symmetry=True)
    player = make_player(q, net)
    player.mcts.seed(0)
    game = Game(mod.Env())
    runs = []
    for rep in range(p.get("reps", 1) + 1):                      # first call = warm-up (cuDNN plans, allocator)
        n_games = p["n_games"] if rep else min(p["n_games"], p.get("warm_games", p["n_games"]))
        if n_games != player.n_envs:
            player = make_player(dict(q, n_games=n_games), net)
        t0 = time.perf_counter()
        done = game.batch_self_play(player, n_games, 1.0, 20, 0.0, td_steps=10)
        if dev != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        plies = int(sum(len(pd) - 1 for _, pd in done))
        runs.append(dict(seconds=dt, games=len(done), plies=plies, sims=plies * p["n_playout"]))
    best = min(runs[1:], key=lambda r: r["seconds"] / r["games"])
    res = dict(games_per_sec=best["games"] / best["seconds"], sims_per_sec=best["sims"] / best["seconds"], runs=runs, device=dev,
               engine_module=type(player.mcts.mcts).__module__, net=type(net).__module__)
    with open(out, "w") as f:
        json.dump(res, f)
    print("actor", json.dumps(res))


if __name__ == "__main__":
    task, out, params = sys.argv[1], sys.argv[2], json.loads(sys.argv[3])
    {"selfplay": task_selfplay, "playout": task_playout, "actor": task_actor, "cnn": task_cnn, "load_pt": task_load_pt}[task](out, params)
