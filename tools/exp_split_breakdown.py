"""Where does one iteration of the split host-buffer API spend its time?  python tools/exp_split_breakdown.py [games]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K, A = 4, 7
boards, turns = bench.c4_random_roots(G, 1000)
eng = mcts_cpp.BatchedMCTS_Connect4(G, device=0)
for k, v in bench.SERVER_DEFAULTS.items():
    setattr(eng.config, k, v)
reset = np.full(G, -1, np.int32)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
def evaluate(lt, it, td, tp1, tp2):
    t = it.astype(bool); p1 = lt == 1
    probs = np.ones((it.shape[0], A), np.float32); probs[t] = 0
    d = np.where(t, td, np.float32(0.25)).astype(np.float32)
    p1w = np.where(t, tp1, np.where(p1, np.float32(0.5), np.float32(0.25))).astype(np.float32)
    p2w = np.where(t, tp2, np.where(p1, np.float32(0.25), np.float32(0.5))).astype(np.float32)
    ml = np.where(t, np.float32(0), np.float32(10)).astype(np.float32)
    return probs, d, p1w, p2w, ml
for rep in range(4):
    if rep == 1:
        T.clear()
    eng.prune_roots(reset)
    lb, td, tp1, tp2, it, lt, vm = eng.search_batch(boards, turns)
    eng.backprop_batch(*evaluate(lt, it, td, tp1, tp2), it)
    for _ in range(50):
        t = time.perf_counter()
        lb, td, tp1, tp2, it, lt, sym, vm = eng.search_batch_vl(K, boards, turns); t = tick("search_batch_vl", t)
        ev = evaluate(lt, it, td, tp1, tp2); t = tick("numpy evaluator", t)
        eng.backprop_batch_vl(K, *ev, it, sym); t = tick("backprop_batch_vl", t)
for k, v in T.items():
    print(f"{k:22s} {v / 150 * 1e3:7.3f} ms/iteration")
print(f"games {G}: {G * 200 / (sum(T.values()) / 3) / 1e6:.1f} M sims/s")
