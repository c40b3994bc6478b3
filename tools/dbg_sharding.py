import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from harness import SERVER_DEFAULTS
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
G, npl, K, plies = 256, 40, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 30
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
sym = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
temp = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
cfg = dict(SERVER_DEFAULTS, use_symmetry=sym, dirichlet_alpha=alpha)
def play(n, base):
    sp = sp_mod.SelfPlay("Connect4", n, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=cfg, temperature=temp, temp_decay_moves=8,
                         td_steps=4, seed=21, uid_base=base, uid_stride=G, out_capacity=6 * n)
    snaps = []
    for p in range(plies):
        sp.ply()
        torch.cuda.synchronize()
        snaps.append((sp.counts.cpu().numpy().copy(), sp.actions.cpu().numpy().copy(), sp.stats.cpu().numpy().copy()))
    return sp.drain(), snaps
w, sw = play(G, 0)
a, sa = play(G // 2, 0)
b, sb = play(G // 2, G // 2)
for p in range(plies):
    c = np.concatenate([sa[p][0], sb[p][0]]); ac = np.concatenate([sa[p][1], sb[p][1]]); st = np.concatenate([sa[p][2], sb[p][2]])
    dc = np.where((c != sw[p][0]).any(1))[0]; da = np.where(ac != sw[p][1])[0]; dst = np.where((st != sw[p][2]).any(1))[0]
    if len(dc) or len(da) or len(dst):
        print("ply", p, "counts differ in slots", dc[:10], "actions differ", da[:10], "stats differ", dst[:10])
        i = (list(dc) + list(da) + list(dst))[0]
        print(" slot", i, "whole counts", sw[p][0][i], "halves", c[i], "act", sw[p][1][i], ac[i])
        print(" stats whole", sw[p][2][i][:14], "halves", st[i][:14])
        break
else:
    print("all plies identical; games", len(w), len(a) + len(b))
