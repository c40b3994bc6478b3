"""Self-play with the C4Net stand-in in the loop: plain evaluator vs device cache vs cache + in-batch dedup.
python tools/exp_selfplay_cnn.py [slots] [plies] [cache_entries]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
nets = importlib.import_module("alphazero-al_b200.nets")
n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 24
entries = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22
for mode in ("cache+dedup", "cache", "plain"):
    torch.manual_seed(0)
    net = nets.C4Net(device="cuda:0")
    sp = sp_mod.SelfPlay("Connect4", n_slots, 200, 4, net, search_cfg=bench.SERVER_DEFAULTS, temperature=1.0, temp_decay_moves=20,
                         temp_endgame=0.0, td_steps=10, seed=0, device=0, out_capacity=4 * n_slots, cache_size=0 if mode == "plain" else entries)
    if mode == "cache":
        sp.evaluator.dedup = False
    sp.engine.reserve(16384)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_pl = plies if mode != "plain" else min(plies, 6)
    for p in range(n_pl):
        t1 = time.perf_counter()
        sp.ply()
        torch.cuda.synchronize()
        if mode != "plain" and (p % 6 == 5 or p < 3):
            st = sp.eval_cache.stats()
            print(f"  {mode} ply {p:2d}: {1e3 * (time.perf_counter() - t1):8.1f} ms  lookups {st['lookups']} hits {st['hits']} dups {st['dups']} "
                  f"net rows {sp.evaluator.net_rows}", flush=True)
    dt = time.perf_counter() - t0
    g = sp.finished()
    print(f"{mode:12s}: {n_slots} slots, {n_pl} plies in {dt:.2f} s -> {n_slots * n_pl / dt:9.0f} positions/s  {n_slots * n_pl * 200 / dt / 1e6:8.2f} M sims/s  "
          f"games finished {g}", flush=True)
    del sp, net
