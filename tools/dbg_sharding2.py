import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from harness import SERVER_DEFAULTS
sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
ds = importlib.import_module("alphazero-al_b200.device_search")
G, npl, K, plies = 256, 40, 4, 30
sync = int(sys.argv[1])
cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3)
def play(n, base):
    sp = sp_mod.SelfPlay("Connect4", n, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=cfg, temperature=1.0, temp_decay_moves=8,
                         td_steps=4, seed=21, uid_base=base, uid_stride=G, out_capacity=6 * n)
    for p in range(plies):
        sp.ply()
        if sync: torch.cuda.synchronize()
    return sp.drain()
w = play(G, 0)
a = play(G // 2, 0); b = play(G // 2, G // 2)
print("games", len(w), len(a), len(b), "positions", w.positions, a.positions, b.positions)
ws = w.sorted_by_uid(); hs = sp_mod.Records.cat([a, b]).sorted_by_uid()
uw, uh = ws.uid.cpu().numpy(), hs.uid.cpu().numpy()
print("uids equal", np.array_equal(uw, uh), "whole max", uw.max(), "halves max", uh.max(), "sorted", (np.diff(uw) > 0).all(), (np.diff(uh) > 0).all())
only_w = sorted(set(uw) - set(uh)); only_h = sorted(set(uh) - set(uw))
print("only whole", only_w[:10], "only halves", only_h[:10])
print("len equal", np.array_equal(ws.length.cpu().numpy(), hs.length.cpu().numpy()) if len(uw) == len(uh) else None)
print("games eq", torch.equal(ws.games, hs.games), "pos eq", torch.equal(ws.pos, hs.pos) if ws.pos.shape == hs.pos.shape else None)
