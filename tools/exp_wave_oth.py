"""Othello: sequential lane-group select vs the staggered one (k_select_ws).  python tools/exp_wave_oth.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_configs as bc
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
oth = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, use_symmetry=True, score_utility_factor=0.15, score_scale=8.0)


def run(n, n_playout, K, lanes, wave, steps=3):
    be = bc.random_roots("Othello", n, 30, 0)
    eng = mcts_cpp.BatchedMCTS_Othello(n)
    for k, v in oth.items():
        setattr(eng.config, k, v)
    eng.set_lanes(lanes)
    eng.set_wave_max(wave)
    eng.reserve(n_playout * 40)
    buf = ds.LeafBuffers(n, n * K, 65, (8, 8), be.device)
    buf.roots = be.states
    ev = ds.SyntheticEvaluator("Othello", "hash")
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
    print(f"Othello N={n:6d} n={n_playout} K={K} backprop lanes={lanes:2d} staggered={'on ' if wave else 'off'}: {ms:8.3f} ms/move  {n * n_playout / ms / 1e6:7.3f} G sims/s", flush=True)


for n in (256, 1024, 4096, 8192, 16384, 32768):
    for lanes in ((16, 8) if n <= 8192 else (8,)):
        run(n, 400, 4, lanes, 0)
        run(n, 400, 4, lanes, 1 << 30)
