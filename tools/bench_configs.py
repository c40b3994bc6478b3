#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (they are parity-test cases, not bench lines): device-resident
loop, CUDA-event timed.  python tools/bench_configs.py"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")


def random_roots(game, n, max_plies, seed):
    """mid-game roots produced on the device: config-2 style hash rollouts stopped after (g mod max_plies) plies"""
    be = env_cpp.BatchedEnv(game, n)
    rng = np.random.default_rng(seed)
    target = torch.from_numpy((np.arange(n) % max_plies).astype(np.int32)).to(be.device)
    for ply in range(max_plies):
        obs = be.observe(boards=False)
        m = obs["masks"].float()
        r = torch.rand(m.shape, device=be.device) * m
        acts = torch.argmax(r, dim=1).int()
        acts = torch.where((target > ply) & (obs["dones"] == 0), acts, torch.full_like(acts, -1))
        # do not play a move that ends the game: peek by stepping a copy
        saved = be.states.clone()
        be.step(acts)
        o2 = be.observe(boards=False, masks=False)
        ended = o2["dones"] != 0
        be.states[ended] = saved[ended]
    return be


def run(game, n, n_playout, K, cfg, mode, steps=3, label="", lanes=0):
    be = random_roots(game, n, 20 if game == "Connect4" else 30, 0)
    A = 7 if game == "Connect4" else 65
    shape = (6, 7) if game == "Connect4" else (8, 8)
    eng = getattr(mcts_cpp, f"BatchedMCTS_{game}")(n)
    for k, v in cfg.items():
        setattr(eng.config, k, v)
    if lanes:
        eng.set_lanes(lanes)
    eng.reserve(n_playout * (8 if game == "Connect4" else 40))
    buf = ds.LeafBuffers(n, n * K, A, shape, be.device)
    buf.roots = be.states
    ev = ds.SyntheticEvaluator(game, mode)
    reset = torch.full((n,), -1, dtype=torch.int32, device=be.device)
    s = torch.cuda.current_stream().cuda_stream

    def step():
        eng.prune_roots_dev(reset.data_ptr(), s)
        ds.playout_device(eng, buf, n_playout, K, ev, s)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": label, "game": game, "n_envs": n, "n_playout": n_playout, "vl_batch": K, "evaluator": mode,
                      "lanes": eng.get_lanes(), "ms_per_move": ms, "sims_per_sec": n * n_playout / (ms * 1e-3)}))


if __name__ == "__main__":
    c4 = dict(c_init=1.4, c_base=4000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, mlh_slope=0.1, mlh_cap=0.2, use_symmetry=True)
    oth = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, use_symmetry=True,
               score_utility_factor=0.15, score_scale=8.0)
    c1 = dict(c4, c_base=1000.0)
    if "--small" in sys.argv:
        for n in (100, 1024, 2048, 4096):
            for ln in (1, 4, 8):
                run("Connect4", n, 200, 4, c1, "hash", steps=5, label=f"C4 N={n} n=200 K=4, {ln} lanes", lanes=ln)
        sys.exit(0)
    run("Connect4", 100, 200, 4, c1, "hash", steps=10, label="config 1 search: C4 N=100 n=200 K=4")
    run("Connect4", 8192, 800, 8, c4, "equivariant", label="config 3: C4 N=8192 n=800 K=8 sym+MLH")
    run("Connect4", 16384, 200, 4, c1, "hash", label="C4 N=16384 n=200 K=4")
    run("Connect4", 32768, 200, 4, c1, "hash", label="C4 N=32768 n=200 K=4")
    run("Connect4", 65536, 800, 8, c4, "equivariant", label="config 3 shape at N=65536")
    run("Othello", 4096, 400, 4, oth, "hash", label="config 4: Othello N=4096 n=400 K=4 score utility")
    run("Othello", 4096, 400, 4, oth, "hash", label="config 4, 8 lanes", lanes=8)
    run("Othello", 32768, 400, 4, oth, "hash", label="config 4 shape at N=32768, 16 lanes", lanes=16)
    run("Othello", 32768, 400, 4, oth, "hash", label="config 4 shape at N=32768, 8 lanes", lanes=8)
    # config 2: env-only lockstep random rollouts, 1M games
    be = env_cpp.BatchedEnv("Connect4", 1_000_000)
    be.random_rollouts(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d, p, _ = be.random_rollouts(0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    plies = int(p.sum().item())
    print(json.dumps({"config": "config 2: C4 env-only 1M games random rollouts", "ms": ms, "plies": plies, "plies_per_sec": plies / (ms * 1e-3),
                      "GB_per_s_algorithmic_38B_per_ply": plies * 38 / (ms * 1e-3) / 1e9}))
