"""The reference's own Python stack, UNMODIFIED, on top of this repository's modules: an overlay
directory whose ``src`` package is made of symlinks to the reference's files plus the three shims a maintainer adds
(``src/azb200`` -> this package, ``src/mcts_cpp.py``, ``src/env_cpp/``).  Build-container only (needs /root/reference; nothing is
copied).  On the CPU this covers what needs no engine instance: ``src.environments.load`` picks up our ``Env`` classes, the
reference's ``Game.play`` + ``NetworkPlayer`` (its CNN on the CPU) play whole games on them, identically to the same code on the
compiled reference ``env_cpp``, and ``src.player`` / ``src.MCTS_cpp`` import against our ``mcts_cpp`` surface."""
import os
import subprocess
import sys
import textwrap

import pytest

import oracle

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")) or not oracle.ref_available("parity"),
                                reason="needs the reference checkout and oracle/_ref/parity (build container only)")

SCRIPT = textwrap.dedent("""
    import sys, numpy as np, torch
    which = sys.argv[1]
    if which == "reference":                       # the compiled, unmodified reference modules under the names src.* expects
        sys.path.insert(0, sys.argv[2])
        import oracle
        mcts_cpp, env_cpp = oracle.load_ref("parity")
        import src
        sys.modules["src.mcts_cpp"], sys.modules["src.env_cpp"] = mcts_cpp, env_cpp
        src.mcts_cpp, src.env_cpp = mcts_cpp, env_cpp
        for g in ("connect4", "othello", "gomoku"):
            sys.modules[f"src.env_cpp.{g}"] = getattr(env_cpp, g)
    from src.environments import load
    from src.game import Game
    from src.player import NetworkPlayer, AlphaZeroPlayer          # imports src.MCTS_cpp -> src.mcts_cpp
    from src import mcts_cpp
    assert mcts_cpp.BatchedMCTS_Connect4.action_size == 7 and mcts_cpp.BatchedMCTS_Othello.board_shape == (8, 8)
    cfg = mcts_cpp.SearchConfig(); cfg.c_init = 1.4; assert abs(cfg.c_init - 1.4) < 1e-6 and cfg.vl_count == 1
    out = []
    for name in ("Connect4", "Othello"):
        mod = load(name)
        torch.manual_seed(0); np.random.seed(0)
        net = mod.CNN(lr=0.0, device="cpu") if "device" in mod.CNN.__init__.__code__.co_varnames else mod.CNN(lr=0.0)
        net.eval()
        env = mod.Env()
        game = Game(env)
        for g in range(2):
            w = game.play(NetworkPlayer(net, deterministic=(g == 0)), NetworkPlayer(net, deterministic=True), show=0)
            out.append((name, int(w), np.asarray(env.board).astype(np.int8).tobytes().hex(), int(env.turn)))
    gm = load("Gomoku")
    e = gm.Env(9, 5)
    rng = np.random.default_rng(0)
    while not e.done():
        mv = e.valid_move(); e.step(mv[int(rng.integers(0, len(mv)))])
    out.append(("Gomoku", int(e.winPlayer()), np.asarray(e.board).astype(np.int8).tobytes().hex(), int(e.turn)))
    print("MODULES", load("Connect4").Env.__module__, load("Gomoku").Env.__module__, getattr(mcts_cpp, "__file__", "?"))
    print(repr(out))
""")


def _overlay(tmp_path):
    src = tmp_path / "src"
    src.mkdir()
    for name in os.listdir(os.path.join(REF, "src")):
        if name not in ("cpp", "__pycache__") and not name.startswith(("mcts_cpp", "env_cpp")):
            os.symlink(os.path.join(REF, "src", name), src / name)
    (src / "__init__.py").write_text("")
    return src


def _run(cwd, *args):
    env = dict(os.environ, PYTHONPATH=str(cwd) + os.pathsep + ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT, *args], cwd=cwd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.strip().splitlines()
    return lines[-1], [l for l in lines if l.startswith("MODULES")][-1]


def test_reference_python_stack_runs_on_our_modules(tmp_path):
    ours = tmp_path / "ours"
    ours.mkdir()
    src = _overlay(ours)
    os.symlink(os.path.join(ROOT, "alphazero-al_b200"), src / "azb200")
    (src / "mcts_cpp.py").write_text("from src.azb200.mcts_cpp import *\n")
    (src / "env_cpp").mkdir()
    (src / "env_cpp" / "__init__.py").write_text("from src.azb200.env_cpp import *\n")
    for g in ("connect4", "othello", "gomoku"):
        (src / "env_cpp" / f"{g}.py").write_text(f"from src.azb200.env_cpp.{g} import *\n")
    theirs = tmp_path / "theirs"
    theirs.mkdir()
    _overlay(theirs)
    (a, mods_a), (b, mods_b) = _run(ours, "ours"), _run(theirs, "reference", ROOT)
    assert mods_a.count("src.azb200.env_cpp.") == 2 and mods_a.endswith("ours/src/mcts_cpp.py") and "azb200" not in mods_b and "oracle/_ref/parity" in mods_b, (mods_a, mods_b)
    assert a == b and a.count("Connect4") == 2 and a.count("Othello") == 2 and "Gomoku" in a
