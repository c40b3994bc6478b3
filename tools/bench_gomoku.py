"""Gomoku lockstep rollouts on the device (az_gomoku_rollout_dev): plies/s at 15x15, five in a row, and the lockstep
step + observe kernels.  CUDA events on the launching stream, after warm-up."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    out = {}
    for n in (65536, 1 << 20):
        be = env_cpp.BatchedGomoku(n, 15, 5)
        plies = [None]

        def roll():
            plies[0] = be.random_rollouts(seed=0, first_game=0, keep_final=False)[1]
        ms = timed(roll)
        total = int(plies[0].sum().item())
        out[f"rollout_n{n}"] = dict(ms=ms, plies=total, plies_per_sec=total / ms * 1e3, games_per_sec=n / ms * 1e3)
        be.random_rollouts(seed=0, first_game=0)                       # finished records in states
        acts = torch.zeros(n, dtype=torch.int32, device=be.device)
        be.reset()
        out[f"step_n{n}"] = dict(ms=timed(lambda: be.step(acts)), record_bytes=288)
        out[f"observe_n{n}"] = dict(ms=timed(lambda: be.observe()), bytes_out=n * 225 * 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
