"""Golden vectors generated from the UNMODIFIED reference engine (tests/golden/make_golden.py, oracle/_ref/parity).
CPU: the C restatement must reproduce them (pins the oracle where oracle/_ref is absent).
GPU: the CUDA engine, through the C ABI, must reproduce them too."""
import importlib
import importlib.util
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def _check(engine, name):
    case = mg.CASES[name]
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    c, s, a, first = mg.run_case(engine, case, z["boards"], z["turns"])
    assert np.array_equal(a, z["actions"])
    assert np.array_equal(c, z["counts"]), "visit counts differ from the reference"
    assert s.tobytes() == z["stats"].tobytes(), "root statistics differ from the reference"
    for j, x in enumerate(first):
        assert np.array_equal(x, z[f"leaf{j}"]), f"leaf output {j} differs from the reference"


@pytest.mark.parametrize("name", list(mg.CASES))
def test_restatement_reproduces_reference_golden(name):
    case = mg.CASES[name]
    _check(oracle.OracleMCTS(case["game"], case["n"]), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_cuda_engine_reproduces_reference_golden(name):
    case = mg.CASES[name]
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    _check(getattr(m, f"BatchedMCTS_{case['game']}")(case["n"]), name)


@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_restatement_env_reproduces_reference_env_games(game):
    z = np.load(os.path.join(HERE, "golden", f"env_{game.lower()}_games.npz"))
    g = 0
    while f"g{g}_actions" in z:
        e = oracle.OracleEnv(game)
        acts = z[f"g{g}_actions"]
        for t, a in enumerate(acts):
            assert np.array_equal(e.board, z[f"g{g}_boards"][t])
            mask = np.zeros(e.A, np.uint8)
            mask[e.valid_moves()] = 1
            assert np.array_equal(mask, z[f"g{g}_masks"][t])
            assert e.turn == z[f"g{g}_turns"][t]
            e.step(int(a))
            assert e.winner() == z[f"g{g}_winners"][t] and e.done() == bool(z[f"g{g}_dones"][t])
        assert np.array_equal(e.board, z[f"g{g}_final"])
        g += 1
    assert g == 40


# ---- Gomoku (Env-only): games and imported boards recorded from the reference's env_cpp.gomoku.Env ----
def _gomoku_games():
    z = np.load(os.path.join(HERE, "golden", "env_gomoku_games.npz"))
    g = 0
    while f"g{g}_actions" in z:
        yield {k: z[f"g{g}_{k}"] for k in ("params", "actions", "boards", "winners", "dones", "turns", "final", "sym")}
        g += 1
    assert g == sum(c[2] for c in mg.GOMOKU_CASES)


def _replay_gomoku(make, board_of, turn_of, step, winner_of, done_of, sym_board):
    for d in _gomoku_games():
        size, k = (int(v) for v in d["params"])
        e = make(size, k)
        for t, a in enumerate(d["actions"]):
            assert np.array_equal(board_of(e), d["boards"][t]) and turn_of(e) == d["turns"][t]
            step(e, int(a))
            assert winner_of(e) == d["winners"][t] and done_of(e) == bool(d["dones"][t])
        assert np.array_equal(board_of(e), d["final"])
        for s in range(8):
            assert np.array_equal(sym_board(e, s, size, k), d["sym"][s])


def test_restatement_gomoku_reproduces_reference_games():
    def sym_board(e, s, size, k):
        e2 = oracle.OracleGomoku(size, k)
        e2.import_board(e.board)
        e2.apply_symmetry(s)
        return e2.board
    _replay_gomoku(oracle.OracleGomoku, lambda e: e.board, lambda e: e.turn, lambda e, a: e.step(a), lambda e: e.winner(),
                   lambda e: e.done(), sym_board)
    z = np.load(os.path.join(HERE, "golden", "env_gomoku_games.npz"))
    for b, (turn, done, winner) in zip(z["import_boards"], z["import_results"]):
        e = oracle.OracleGomoku(8, 4)
        assert e.import_board(b) == 0 and (e.turn, int(e.done()), e.winner()) == (turn, done, winner)


def test_host_gomoku_env_reproduces_reference_games():
    gomoku = importlib.import_module("alphazero-al_b200.env_cpp.gomoku")
    _replay_gomoku(gomoku.Env, lambda e: e.board.astype(np.int8), lambda e: e.turn, lambda e, a: e.step(a), lambda e: e.winPlayer(),
                   lambda e: e.done(), lambda e, s, size, k: e.apply_symmetry(s).board.astype(np.int8))
    z = np.load(os.path.join(HERE, "golden", "env_gomoku_games.npz"))
    for b, (turn, done, winner) in zip(z["import_boards"], z["import_results"]):
        e = gomoku.Env(b.astype(np.float32), 4)
        assert (e.turn, int(e.done()), e.winPlayer()) == (turn, done, winner)


@pytest.mark.gpu
def test_cuda_gomoku_lockstep_reproduces_reference_games():
    """All recorded games of one board size advance in lockstep on the device (shorter games idle with action -1)."""
    import torch
    env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")
    games = list(_gomoku_games())
    for size, k, _ in mg.GOMOKU_CASES:
        gs = [d for d in games if tuple(int(v) for v in d["params"]) == (size, k)]
        be = env_cpp.BatchedGomoku(len(gs), size, k)
        status = torch.zeros(len(gs), dtype=torch.uint8, device=be.device)
        for t in range(max(len(d["actions"]) for d in gs)):
            obs = {kk: v.cpu().numpy() for kk, v in be.observe().items()}
            acts = np.full(len(gs), -1, np.int32)
            for i, d in enumerate(gs):
                if t < len(d["actions"]):
                    assert np.array_equal(obs["boards"][i], d["boards"][t]) and obs["turns"][i] == d["turns"][t]
                    assert np.array_equal(obs["masks"][i], (d["boards"][t].reshape(-1) == 0).astype(np.uint8))
                    acts[i] = d["actions"][t]
                else:
                    assert np.array_equal(obs["boards"][i], d["final"]) and obs["dones"][i] == 1
                    assert obs["winners"][i] == d["winners"][-1]
            be.step(torch.from_numpy(acts).to(be.device), status)
            assert not status.cpu().numpy().any()
        obs = {kk: v.cpu().numpy() for kk, v in be.observe().items()}
        for i, d in enumerate(gs):
            assert np.array_equal(obs["boards"][i], d["final"]) and obs["winners"][i] == d["winners"][-1] and obs["dones"][i] == 1
        for s in range(8):
            bs = env_cpp.BatchedGomoku(len(gs), size, k)
            bs.states.copy_(be.states)
            bs.apply_symmetry(torch.full((len(gs),), s, dtype=torch.int32, device=be.device))
            sb = bs.observe()["boards"].cpu().numpy()
            for i, d in enumerate(gs):
                assert np.array_equal(sb[i], d["sym"][s])
