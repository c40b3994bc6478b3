"""Experiment: split the batch into S shards driven on S streams so that select of one shard overlaps evaluate/backprop of
another (both kernels are latency bound with idle issue slots).  python tools/exp_pipeline.py"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
G, n_playout, K = 65536, 200, 4
dev = torch.device("cuda", 0)
boards_np, turns_np = bench.c4_random_roots(G, 1000)
for S in (1, 2, 4):
    n = G // S
    engs, bufs, streams, resets = [], [], [], []
    for s in range(S):
        e = mcts_cpp.BatchedMCTS_Connect4(n, device=0)
        for k, v in bench.SERVER_DEFAULTS.items():
            setattr(e.config, k, v)
        e.set_lanes(1)
        b = ds.LeafBuffers(n, n * K, 7, (6, 7), dev)
        st = torch.cuda.Stream()
        b.pack_roots(torch.from_numpy(boards_np[s * n:(s + 1) * n]).to(dev), torch.from_numpy(turns_np[s * n:(s + 1) * n]).to(dev), 0)
        engs.append(e); bufs.append(b); streams.append(st); resets.append(torch.full((n,), -1, dtype=torch.int32, device=dev))
    ev = ds.SyntheticEvaluator("Connect4", "constant")
    torch.cuda.synchronize()

    def step():
        for s in range(S):
            engs[s].prune_roots_dev(resets[s].data_ptr(), streams[s].cuda_stream)
        ks = [0] + [min(K, r) for r in [K] * ((n_playout - 1) // K)] + ([(n_playout - 1) % K] if (n_playout - 1) % K else [])
        for k in ks:
            for s in range(S):
                cs = streams[s].cuda_stream
                rows = n * max(k, 1)
                engs[s].search_dev(k, bufs[s].roots.data_ptr(), bufs[s].leaves.data_ptr(), cs)
                ev(bufs[s], rows, cs)
                engs[s].backprop_dev(k, bufs[s].policy.data_ptr(), bufs[s].d.data_ptr(), bufs[s].p1w.data_ptr(), bufs[s].p2w.data_ptr(),
                                     bufs[s].ml.data_ptr(), 0, 0, cs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    c = np.concatenate([e.get_all_counts_array() for e in engs])
    assert (c.sum(1) == n_playout - 1).all()
    print(f"shards {S}: {dt*1e3:.2f} ms/step  {G*n_playout/dt/1e6:.0f} Msims/s")
    del engs, bufs
