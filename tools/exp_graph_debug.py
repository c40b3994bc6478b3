import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import SERVER_DEFAULTS, random_positions, set_config, counts
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
n, K, npl = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 4, 61
cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.0)
boards, turns = random_positions("Connect4", 64, 14, 41)
boards, turns = np.tile(boards, ((n + 63) // 64, 1, 1))[:n], np.tile(turns, (n + 63) // 64)[:n]
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
res = {}
for graphs in ("0", "1"):
    for sh in (1, 2, 4):
        for calls in (1,):
            os.environ["AZB200_GRAPHS"] = graphs
            e = mcts_cpp.BatchedMCTS_Connect4(n)
            e.set_lanes(1); set_config(e, **cfg); e.set_seed(5)
            buf = ds.LeafBuffers(n, n * K, 7, (6, 7), dev)
            buf.pack_roots(torch.from_numpy(boards).to(dev), torch.from_numpy(turns).to(dev), stream)
            for _ in range(calls):
                ds.playout_device(e, buf, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), stream, shards=sh)
            res[(graphs, sh, calls)] = counts(e, n, 7)
ref = {c: res[("0", 1, c)] for c in (1,)}
for k, v in res.items():
    bad = np.where((v != ref[k[2]]).any(axis=1))[0]
    print(k, "mismatching trees:", len(bad), bad[:10])
