"""Where does the wrapper-level (host-buffer) step spend its time?  python tools/exp_e2e_breakdown.py"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
bm = importlib.import_module("alphazero-al_b200.batched_mcts")
ds = importlib.import_module("alphazero-al_b200.device_search")
G, n_playout, K = 65536, 200, 4
D = bench.SERVER_DEFAULTS
boards_np, turns_np = bench.c4_random_roots(G, 1000)
reset_np = np.full(G, -1, np.int32)
ev = ds.SyntheticEvaluator("Connect4", "constant")
wrap = bm.BatchedMCTS(G, D["c_init"], D["c_base"], D["dirichlet_alpha"], n_playout, game_name="Connect4", noise_epsilon=D["noise_epsilon"],
                      fpu_reduction=D["fpu_reduction"], use_symmetry=True, mlh_slope=D["mlh_slope"], mlh_cap=D["mlh_cap"], device=0)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
for it in range(8):
    if it == 3:
        T.clear()
    t = time.perf_counter()
    wrap.prune_roots(reset_np); t = tick("prune_roots", t)
    wrap.batch_playout(ev, boards_np, turns_np, vl_batch=K); t = tick("batch_playout (enqueue)", t)
    torch.cuda.synchronize(); t = tick("wait for GPU", t)
    c = wrap.get_visits_count(); t = tick("get_visits_count", t)
for k, v in T.items():
    print(f"{k:28s} {v / 5 * 1e3:7.3f} ms/step")
print("total", sum(T.values()) / 5 * 1e3)
