"""One steady-state self-play ply for an ncu launch list:  ncu --metrics gpu__time_duration.sum --launch-skip N ... python tools/exp_selfplay_ncu.py [slots] [plies]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 26
sp = sp_mod.SelfPlay("Connect4", n_slots, 200, 4, ds.SyntheticEvaluator("Connect4", "constant"), search_cfg=bench.SERVER_DEFAULTS,
                     temperature=1.0, temp_decay_moves=20, temp_endgame=0.0, td_steps=10, seed=0, device=0, out_capacity=4 * n_slots)
sp.engine.reserve(8192)
for _ in range(plies):
    sp.ply()
torch.cuda.synchronize()
print("launches per ply", sp.launches / plies, "total", sp.launches)
