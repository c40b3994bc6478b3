"""Staggered-descent select (az_mcts_wave.cuh) against the thread-per-tree and 8-lane kernels on small batches.
python tools/exp_wave.py"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_configs as bc
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")


def run(n, n_playout, K, cfg, mode, lanes, wave, steps=3):
    be = bc.random_roots("Connect4", n, 20, 0)
    eng = mcts_cpp.BatchedMCTS_Connect4(n)
    for k, v in cfg.items():
        setattr(eng.config, k, v)
    eng.set_lanes(lanes)
    eng.set_wave_max(wave)
    eng.reserve(n_playout * 8)
    buf = ds.LeafBuffers(n, n * K, 7, (6, 7), be.device)
    buf.roots = be.states
    ev = ds.SyntheticEvaluator("Connect4", mode)
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
    print(f"N={n:6d} n={n_playout} K={K} lanes={lanes} wave={'on ' if wave else 'off'}: {ms:8.3f} ms/move  {n * n_playout / ms / 1e6:9.2f} M sims/s", flush=True)


c4 = dict(c_init=1.4, c_base=4000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, mlh_slope=0.1, mlh_cap=0.2, use_symmetry=True)
c1 = dict(c4, c_base=1000.0)
for n in (100, 1024, 4096):
    run(n, 200, 4, c1, "hash", 8, 0)
    run(n, 200, 4, c1, "hash", 1, 0)
    run(n, 200, 4, c1, "hash", 1, 1 << 20)
for n in (8192, 16384, 32768):
    run(n, 200, 4, c1, "hash", 1, 0)
    run(n, 200, 4, c1, "hash", 1, 1 << 20)
for n in (2048, 8192, 16384, 32768):
    run(n, 800, 8, c4, "equivariant", 1, 0)
    run(n, 800, 8, c4, "equivariant", 1, 1 << 20)
