"""Generates tests/golden/*.npz from the UNMODIFIED reference engine (oracle/_ref/parity, built by `make -C oracle ref`
from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture stores inputs (root boards, turns, config) and the reference's outputs (visit counts, root statistics
after every move, first-iteration leaves) under the deterministic hash evaluator of alphazero-al_b200/evaluators.py."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from harness import SERVER_DEFAULTS, counts, playout, random_positions, set_config  # noqa: E402

ev_mod = importlib.import_module("alphazero-al_b200.evaluators")

CASES = {
    "c4_n200_k4_reuse": dict(game="Connect4", n=32, n_playout=200, K=4, moves=4, max_plies=18, seed=101, mode="hash",
                             cfg=SERVER_DEFAULTS),
    "c4_n64_k8_decay": dict(game="Connect4", n=32, n_playout=64, K=8, moves=6, max_plies=36, seed=102, mode="hash",
                            cfg=dict(SERVER_DEFAULTS, value_decay=0.97, mlh_slope=0.0)),
    "c4_nonvl_default_cfg": dict(game="Connect4", n=16, n_playout=50, K=1, moves=2, max_plies=10, seed=103, mode="hash",
                                 cfg=dict(dirichlet_alpha=0.0, use_symmetry=False)),
    "oth_n120_k4_score": dict(game="Othello", n=16, n_playout=120, K=4, moves=4, max_plies=50, seed=104, mode="hash",
                              cfg=dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False,
                                       score_utility_factor=0.15, score_scale=8.0)),
}


def run_case(engine, spec, boards, turns):
    """Returns per-move (counts, root_stats, actions) and the leaves of the first iteration of the first move."""
    game, n, A = spec["game"], spec["n"], oracle.ACTION_SIZE[spec["game"]]
    set_config(engine, **spec["cfg"])
    ev = ev_mod.HashEvaluator(game, spec["mode"])
    envs = [oracle.OracleEnv(game) for _ in range(n)]
    for i, e in enumerate(envs):
        e.import_board(boards[i], turns[i])
    out_counts, out_stats, out_actions, first = [], [], [], None
    for mv in range(spec["moves"]):
        b = np.stack([e.board for e in envs])
        t = np.array([e.turn for e in envs], np.int32)
        rec = []
        playout(engine, ev, b, t, spec["n_playout"], spec["K"], rec)
        if first is None:
            first = rec[1] if len(rec) > 1 else rec[0]
        c = counts(engine, n, A)
        out_counts.append(c)
        out_stats.append(engine.get_all_root_stats().copy())
        acts = np.zeros(n, np.int32)
        for i, e in enumerate(envs):
            if e.done() or c[i].sum() == 0:
                e.reset()
                engine.reset_env(i)
                acts[i] = -1
            else:
                acts[i] = int(np.argmax(c[i]))
                e.step(acts[i])
        out_actions.append(acts)
        engine.prune_roots(acts)
    return np.stack(out_counts), np.stack(out_stats), np.stack(out_actions), first


def main():
    mcts_cpp, env_cpp = oracle.load_ref("parity")
    for name, spec in CASES.items():
        boards, turns = random_positions(spec["game"], spec["n"], spec["max_plies"], spec["seed"])
        eng = getattr(mcts_cpp, f"BatchedMCTS_{spec['game']}")(spec["n"])
        c, s, a, first = run_case(eng, spec, boards, turns)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), boards=boards, turns=turns, counts=c, stats=s, actions=a,
                            **{f"leaf{j}": x for j, x in enumerate(first)})
        print(name, c.shape, s.shape)
    # env play-outs from the reference Env objects: per-ply boards / masks / winner / done for fixed action scripts
    rng = np.random.default_rng(7)
    for game, sub in (("Connect4", env_cpp.connect4), ("Othello", env_cpp.othello)):
        games = []
        for g in range(40):
            e = sub.Env()
            acts, boards_, masks, winners, dones, turns_ = [], [], [], [], [], []
            while not e.done():
                mv = e.valid_move()
                a = mv[int(rng.integers(0, len(mv)))]
                boards_.append(np.asarray(e.board).astype(np.int8))
                masks.append(np.asarray(e.valid_mask(), dtype=np.uint8))
                turns_.append(e.turn)
                e.step(a)
                acts.append(a); winners.append(e.winPlayer()); dones.append(e.done())
            games.append(dict(actions=np.array(acts, np.int32), boards=np.stack(boards_), masks=np.stack(masks),
                              winners=np.array(winners, np.int32), dones=np.array(dones, np.uint8), turns=np.array(turns_, np.int32),
                              final=np.asarray(e.board).astype(np.int8)))
        np.savez_compressed(os.path.join(HERE, f"env_{game.lower()}_games.npz"),
                            **{f"g{i}_{k}": v for i, d in enumerate(games) for k, v in d.items()})
        print("env", game, len(games))


GOMOKU_CASES = ((15, 5, 10), (9, 4, 10), (6, 3, 10), (19, 6, 2))


def gomoku_main():
    """tests/golden/env_gomoku_games.npz from the reference's env_cpp.gomoku.Env: random games (per-ply boards, turns, winner /
    done after each move, final board) and unplayable imported boards with the turn / done / winner the reference infers."""
    _, env_cpp = oracle.load_ref("parity")
    rng = np.random.default_rng(17)
    out, gi = {}, 0
    for size, k, games in GOMOKU_CASES:
        for _ in range(games):
            e = env_cpp.gomoku.Env(size, k)
            acts, boards_, winners, dones, turns_ = [], [], [], [], []
            while not e.done():
                mv = e.valid_move()
                a = mv[int(rng.integers(0, len(mv)))]
                boards_.append(np.asarray(e.board).astype(np.int8))
                turns_.append(e.turn)
                e.step(a)
                acts.append(a); winners.append(e.winPlayer()); dones.append(e.done())
            d = dict(params=np.array([size, k], np.int32), actions=np.array(acts, np.int32), boards=np.stack(boards_),
                     winners=np.array(winners, np.int32), dones=np.array(dones, np.uint8), turns=np.array(turns_, np.int32),
                     final=np.asarray(e.board).astype(np.int8),
                     sym=np.stack([np.asarray(e.apply_symmetry(s).board).astype(np.int8) for s in range(8)]))
            out.update({f"g{gi}_{kk}": v for kk, v in d.items()})
            gi += 1
    imp_boards, imp_res = [], []
    for j in range(120):
        b = rng.choice(np.array([-1, 0, 1], np.int8), size=(8, 8), p=(0.3 + 0.1 * (j % 3), 0.3, 0.4 - 0.1 * (j % 3)))
        e = env_cpp.gomoku.Env(b.astype(np.float32), 4)
        imp_boards.append(b)
        imp_res.append((e.turn, int(e.done()), e.winPlayer()))
    out["import_boards"], out["import_results"] = np.stack(imp_boards), np.array(imp_res, np.int32)
    np.savez_compressed(os.path.join(HERE, "env_gomoku_games.npz"), **out)
    print("env Gomoku", gi, "games,", len(imp_boards), "imported boards")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gomoku":
        gomoku_main()
    else:
        main()
        gomoku_main()
