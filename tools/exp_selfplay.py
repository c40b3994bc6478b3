"""Self-play leg timing per ply (continuous self-play with tree reuse, constant evaluator).  python tools/exp_selfplay.py [slots]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reserve = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sp = sp_mod.SelfPlay("Connect4", n_slots, 200, 4, ds.SyntheticEvaluator("Connect4", "constant"), search_cfg=bench.SERVER_DEFAULTS,
                     temperature=1.0, temp_decay_moves=20, temp_endgame=0.0, td_steps=10, seed=0, device=0, out_capacity=4 * n_slots)
sp.engine.reserve(reserve)
torch.cuda.synchronize()
g_prev = 0
t_tot, g_start = 0.0, 0
for ply in range(48):
    t0 = time.perf_counter()
    sp.ply()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g = sp.finished()
    st = sp.engine.get_stats()
    if ply % 4 == 3 or dt > 0.05:
        print(f"ply {ply:2d}: {dt*1e3:7.2f} ms  games done {g:6d} (+{g-g_prev})  arena cap {st['arena_cap']} max used {st['max_arena_slots']} lanes {sp.engine.get_lanes()}")
    g_prev = g
    if ply >= 24:
        t_tot += dt
    else:
        g_start = g
print(f'plies 24-47: {t_tot / 24 * 1e3:.2f} ms/ply  {n_slots * 200 * 24 / t_tot / 1e9:.3f} G sims/s  {(g - g_start) / t_tot / 1e3:.1f} k games/s  compactions {sp.engine.compactions()}')
