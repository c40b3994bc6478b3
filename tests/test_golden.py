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
