"""Where does a self-play ply go?  CUDA events around the phases of SelfPlay.ply at the bench's size (65 536 slots, n=200, K=4,
constant evaluator, steady state).  python tools/exp_selfplay_breakdown.py [slots]"""
import ctypes as C, importlib, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sp = sp_mod.SelfPlay("Connect4", n, 200, 4, ds.SyntheticEvaluator("Connect4", "constant"), search_cfg=bench.SERVER_DEFAULTS, temperature=1.0,
                     temp_decay_moves=20, td_steps=10, seed=0, out_capacity=2 * n)
sp.engine.reserve(16384)
for _ in range(14):
    sp.ply()
sp.drain()
torch.cuda.synchronize()
L = sp._L
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {"playout": 0.0, "counts+stats": 0.0, "ply": 0.0, "prune(+compaction)": 0.0, "flush": 0.0}
host = 0.0
P = 30
c0 = sp.engine.compactions()
st0 = None
t_all0 = time.perf_counter()
for _ in range(P):
    s = sp._stream()
    e = [ev() for _ in range(6)]
    e[0].record()
    h0 = time.perf_counter()
    ds.playout_device(sp.engine, sp.buf, sp.n_playout, sp.K, sp.evaluator, s or 0)
    e[1].record()
    sp.engine.get_counts_dev(sp.counts.data_ptr(), s or 0); sp.engine.get_root_stats_dev(sp.stats.data_ptr(), s or 0)
    e[2].record()
    L.az_selfplay_ply_dev(C.byref(sp.sp), sp.counts.data_ptr(), sp.stats.data_ptr(), s)
    e[3].record()
    sp.engine.prune_roots_dev(sp.actions.data_ptr(), s or 0)
    e[4].record()
    L.az_selfplay_flush_dev(C.byref(sp.sp), s)
    e[5].record()
    host += time.perf_counter() - h0
    torch.cuda.synchronize()
    for k, (a, b) in zip(acc, zip(e[:-1], e[1:])):
        acc[k] += a.elapsed_time(b)
    sp.plies += 1
wall = (time.perf_counter() - t_all0) / P * 1e3
sp.engine.enable_stats(True)
ds.playout_device(sp.engine, sp.buf, sp.n_playout, sp.K, sp.evaluator, sp._stream() or 0)
torch.cuda.synchronize()
st = sp.engine.get_stats()
print(json.dumps({"slots": n, "ms_per_ply_phase": {k: v / P for k, v in acc.items()}, "host_enqueue_ms_per_ply": host / P * 1e3, "wall_ms_per_ply_with_sync": wall,
                  "compactions_in_%d_plies" % P: sp.engine.compactions() - c0,
                  "tree_stats_one_playout": {"depth": st["depth"] / st["sims"], "edges_scanned": st["edges_scanned"] / st["sims"], "edges_created": st["edges_created"] / st["sims"],
                                             "max_arena_slots": st["max_arena_slots"], "arena_cap": st["arena_cap"]}}))
