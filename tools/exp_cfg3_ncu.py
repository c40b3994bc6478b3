"""One move of BASELINE config 3 (8192 trees, n=800, K=8) for an ncu launch list.  python tools/exp_cfg3_ncu.py [n] [n_playout] [K]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_configs as bc
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
npl = int(sys.argv[2]) if len(sys.argv) > 2 else 800
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cfg = dict(c_init=1.4, c_base=4000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, mlh_slope=0.1, mlh_cap=0.2, use_symmetry=True)
be = bc.random_roots("Connect4", n, 20, 0)
eng = mcts_cpp.BatchedMCTS_Connect4(n)
for k, v in cfg.items():
    setattr(eng.config, k, v)
eng.reserve(npl * 8)
buf = ds.LeafBuffers(n, n * K, 7, (6, 7), be.device)
buf.roots = be.states
ev = ds.SyntheticEvaluator("Connect4", "equivariant")
reset = torch.full((n,), -1, dtype=torch.int32, device=be.device)
s = torch.cuda.current_stream().cuda_stream
eng.prune_roots_dev(reset.data_ptr(), s)
ds.playout_device(eng, buf, npl, K, ev, s)
torch.cuda.synchronize()
# select share of the move in the running pipeline (warm caches): CUDA events around every select launch
if os.environ.get("AZB200_TIME_SELECT"):
    for wave in (0, eng.get_wave_max()):
        eng.set_wave_max(wave)
        for rep in range(2):
            eng.prune_roots_dev(reset.data_ptr(), s)
            torch.cuda.synchronize()
            if rep == 1:
                eng.time_select(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ds.playout_device(eng, buf, npl, K, ev, s)
            e1.record()
            torch.cuda.synchronize()
        ms, nl, rows = eng.get_select_time()
        eng.time_select(False)
        print(f"N={n} n={npl} K={K} wave={'on' if wave else 'off'}: move {e0.elapsed_time(e1):.3f} ms (launch by launch), select {ms:.3f} ms in {nl} launches = {1e3 * ms / max(nl, 1):.1f} us each")
