"""One search step of the bench workload (Connect4, n=200, K=4, 65 536 fresh mid-game roots, constant evaluator) inside a
cudaProfilerStart/Stop range, for ncu (--profile-from-start off).  Whole batch per launch on one stream (the launch shape the
bench's `roofline` block times), after warm-up steps in the bench's own sharded form so that L2 and the allocator are in their
steady state.  python tools/traffic_probe.py [games] [shards_in_range] [game]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shards_in_range = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
boards_np, turns_np = bench.c4_random_roots(G, seed=1000)
boards, turns = torch.from_numpy(boards_np).to(dev), torch.from_numpy(turns_np).to(dev)
eng = mcts_cpp.BatchedMCTS_Connect4(G, device=0)
for k, v in bench.SERVER_DEFAULTS.items():
    setattr(eng.config, k, v)
eng.set_seed(0)
buf = ds.LeafBuffers(G, G * 4, 7, (6, 7), dev)
ev = ds.SyntheticEvaluator("Connect4", "constant")
stream = torch.cuda.current_stream().cuda_stream


def step(shards):
    eng.reset_all_dev(stream)
    buf.pack_roots(boards, turns, stream)
    ds.playout_device(eng, buf, 200, 4, ev, stream, shards=shards)


for _ in range(3):
    step(ds.auto_shards(G))
step(shards_in_range)                       # builds the graph of the profiled launch shape outside the range
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(shards_in_range)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("probe done")
