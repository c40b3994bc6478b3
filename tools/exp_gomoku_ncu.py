"""One launch each of the Gomoku lockstep kernels at 1 M games of 15 x 15 (for ncu):
    ncu --set full --clock-control none -k regex:k_gmk_ -o gpurun_out/gomoku python tools/exp_gomoku_ncu.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
be = env_cpp.BatchedGomoku(n, 15, 5)
be.random_rollouts(seed=0, first_game=0)                      # k_gmk_rollout (final records -> states)
be.observe()                                                  # k_gmk_observe on finished games
be.reset()
acts = torch.full((n,), 112, dtype=torch.int32, device=be.device)
be.step(acts)                                                 # k_gmk_step
torch.cuda.synchronize()
print("ok")
