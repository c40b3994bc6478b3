"""Manual GPU smoke/debug driver (not collected by pytest): python tests/gpu_debug.py"""
import importlib, sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from harness import SERVER_DEFAULTS, compare_engines, random_positions

m = importlib.import_module("alphazero-al_b200.mcts_cpp")
for game, n, npl, K, cfg in (("Connect4", 8, 30, 1, SERVER_DEFAULTS), ("Connect4", 8, 30, 4, SERVER_DEFAULTS),
                             ("Othello", 8, 30, 4, dict(dirichlet_alpha=0.0, use_symmetry=False))):
    try:
        t0 = time.time()
        compare_engines(getattr(m, f"BatchedMCTS_{game}")(n), oracle.OracleMCTS(game, n), game, n, npl, K, cfg, moves=2)
        print(game, K, "OK", round(time.time() - t0, 2), "s")
    except Exception:
        traceback.print_exc()
