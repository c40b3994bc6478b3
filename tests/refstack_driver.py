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
    {"selfplay": task_selfplay, "playout": task_playout, "actor": task_actor}[task](out, params)
