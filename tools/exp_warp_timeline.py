"""How evenly do the warps of one select launch finish?  (per-warp %globaltimer stamps, stats mode)
python tools/exp_warp_timeline.py"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
ds = importlib.import_module("alphazero-al_b200.device_search")
L = importlib.import_module("alphazero-al_b200._lib").lib()
G, K = 65536, 4
dev = torch.device("cuda", 0)
b, t = bench.c4_random_roots(G, 1000)
e = mcts_cpp.BatchedMCTS_Connect4(G, device=0)
for k, v in bench.SERVER_DEFAULTS.items():
    setattr(e.config, k, v)
buf = ds.LeafBuffers(G, G * K, 7, (6, 7), dev)
ev = ds.SyntheticEvaluator("Connect4", "constant")
s = torch.cuda.current_stream().cuda_stream
buf.pack_roots(torch.from_numpy(b).to(dev), torch.from_numpy(t).to(dev), s)
for n_done in (41, 121):                         # look at the select of iteration 10 and 30 of a fresh search
    ds.playout_device(e, buf, 40, K, ev, s)      # 40 more simulations
    e.enable_stats(True)
    e.search_dev(K, buf.roots.data_ptr(), buf.leaves.data_ptr(), s)
    torch.cuda.synchronize()
    out = np.zeros(2 * 2048, np.uint64)
    nw = L.az_mcts_get_warp_times(e._h, out.ctypes.data_as(C.c_void_p), 2048)
    e.enable_stats(False)
    ev(buf, G * K, s)
    e.backprop_dev(K, buf.policy.data_ptr(), buf.d.data_ptr(), buf.p1w.data_ptr(), buf.p2w.data_ptr(), buf.ml.data_ptr(), 0, 0, s)
    raw_en = out[1:2 * nw:2]
    smid, levels = ((raw_en >> np.uint64(56)) & np.uint64(0xFF)).astype(int), ((raw_en >> np.uint64(48)) & np.uint64(0xFF)).astype(int)
    st, en = (out[0:2 * nw:2] & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64), (raw_en & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
    t0 = st.min()
    dur = (en - st) / 1e3
    print(f"after {n_done} sims: warps {nw}  kernel span {(en.max() - t0) / 1e3:.1f} us  start spread {(st.max() - t0) / 1e3:.1f} us  "
          f"warp duration mean {dur.mean():.1f} p50 {np.percentile(dur, 50):.1f} p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f} us")
    endt = (en - t0) / 1e3
    print("   end-time percentiles (us): " + " ".join(f"p{p}={np.percentile(endt, p):.1f}" for p in (10, 50, 90, 99, 100)))
    print(f"   level iterations per warp: mean {levels.mean():.1f} min {levels.min()} max {levels.max()}  corr(duration, levels) = {np.corrcoef(dur, levels)[0, 1]:.2f}")
    by_sm = np.array([dur[smid == i].mean() for i in range(148) if (smid == i).any()])
    cnt = np.array([(smid == i).sum() for i in range(148)])
    print(f"   mean warp duration by SM: min {by_sm.min():.1f} p50 {np.median(by_sm):.1f} max {by_sm.max():.1f} us; warps per SM min {cnt.min()} max {cnt.max()}")
    print("   us per level iteration: mean %.2f p10 %.2f p90 %.2f" % ((dur / levels).mean(), np.percentile(dur / levels, 10), np.percentile(dur / levels, 90)))
    stt = e.get_stats() if hasattr(e, "get_stats") else None
    if stt:
        print("   stats:", {k: stt[k] for k in ("sims", "depth", "edges_scanned")}, "mean depth per simulation %.2f -> %.1f level steps per tree and launch" % (stt["depth"] / max(stt["sims"], 1), K * stt["depth"] / max(stt["sims"], 1)))
