#!/usr/bin/env python
"""Benchmark of the self-play hot path: batched PUCT MCTS simulations/s, Connect4, n=200, vl_batch=4.

    python bench.py [--gpus N --steps K --warmup W]              # this repo (CUDA engine)
    python bench.py --impl reference [...]                        # the reference C++/OpenMP engine on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = one full move search for every game of the batch: fresh trees on G mid-game Connect4 roots per GPU
(config-2 style random rollouts, ply = g mod 20), then n=200 simulations per tree exactly as the reference wrapper
schedules them (src/MCTS_cpp.py:217-357: 1 non-VL warm-up simulation + 50 virtual-loss iterations of K=4/3), with the
server-default search parameters (server.py:44-72) and a deterministic constant evaluator standing in for the
random-init CNN (the CNN is outside the path: SURVEY.md 8d).  Simulations/s = G*n/step time, whole job over all GPUs.

  value : device-resident loop (inputs in HBM, CUDA-event timed, max over ranks; no host synchronisation inside the region)
  e2e   : the same search through the wrapper API with HOST buffers (BatchedMCTS.prune_roots + batch_playout(numpy boards)
          + get_visits_count(), the calls src/player.py makes; H2D/D2H inside the timed region); e2e_split_api: the split
          plugin API (mcts_cpp.search_batch[_vl] / backprop_batch[_vl]) with numpy leaf buffers every iteration
  roofline     : the two tree kernels (select, back-prop): algorithmic bytes / CUDA-event time of their launches vs the measured
                 HBM peak; `traffic` = measured DRAM bytes per launch from the committed ncu summary of THESE kernel sources
  selfplay     : the on-device self-play driver, games/s from finished-game counts, with the trajectory all-gather (the one
                 collective of the path) INSIDE the timed loop at N > 1, overlapped with the next batch's search; plus the
                 strong-scaling form of BASELINE config 5 (65 536 games in total over the N GPUs)
  cpu_baseline : the reference engine (oracle/_ref/timing) on this box's host cores, bounded sample; and the reference ACTOR
                 (its unmodified Game.batch_self_play + its CNN's GPU predict, BASELINE config 1) on this box
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SERVER_DEFAULTS = dict(c_init=1.4, c_base=1000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25,
                       mlh_slope=0.1, mlh_cap=0.2, use_symmetry=True, value_decay=1.0)
METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def c4_random_roots(n: int, seed: int, max_ply: int = 20):
    """n non-terminal Connect4 positions after (g mod max_ply) uniformly random legal plies (numpy, lockstep)."""
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 6, 7), np.int8)
    turns = np.ones(n, np.int32)
    heights = np.zeros((n, 7), np.int64)
    target = np.arange(n) % max_ply

    def four(m):
        h = m[:, :, 0:4] & m[:, :, 1:5] & m[:, :, 2:6] & m[:, :, 3:7]
        v = m[:, 0:3, :] & m[:, 1:4, :] & m[:, 2:5, :] & m[:, 3:6, :]
        d1 = m[:, 0:3, 0:4] & m[:, 1:4, 1:5] & m[:, 2:5, 2:6] & m[:, 3:6, 3:7]
        d2 = m[:, 3:6, 0:4] & m[:, 2:5, 1:5] & m[:, 1:4, 2:6] & m[:, 0:3, 3:7]
        return h.any(axis=(1, 2)) | v.any(axis=(1, 2)) | d1.any(axis=(1, 2)) | d2.any(axis=(1, 2))

    for ply in range(max_ply):
        act = np.where(target > ply)[0]
        if act.size == 0:
            break
        legal = heights[act] < 6
        col = np.argmax(np.where(legal, rng.random((act.size, 7)), -1.0), axis=1)
        row = 5 - heights[act, col]
        trial = boards[act].copy()
        trial[np.arange(act.size), row, col] = turns[act]
        ok = ~four(trial == turns[act][:, None, None])
        good = act[ok]
        boards[good] = trial[ok]
        heights[good, col[ok]] += 1
        turns[good] = -turns[good]
        target[act[~ok]] = ply          # a winning move is not played: the game stays at this ply
    return boards, turns


def host_step(engine, boards, turns, n_playout, K, A, reset_actions):
    """One step through the reference-facing host API with numpy buffers (what src/MCTS_cpp.py does)."""
    engine.prune_roots(reset_actions)               # action -1 matches no edge -> every tree is reset (MCTS.h:107)
    def evaluate(lt, it, td, tp1, tp2):
        t = it.astype(bool)
        p1 = lt == 1
        probs = np.ones((it.shape[0], A), np.float32)
        probs[t] = 0
        d = np.where(t, td, np.float32(0.25)).astype(np.float32)
        p1w = np.where(t, tp1, np.where(p1, np.float32(0.5), np.float32(0.25))).astype(np.float32)
        p2w = np.where(t, tp2, np.where(p1, np.float32(0.25), np.float32(0.5))).astype(np.float32)
        ml = np.where(t, np.float32(0), np.float32(10)).astype(np.float32)
        return probs, d, p1w, p2w, ml
    t_eval, pc = 0.0, time.perf_counter
    lb, td, tp1, tp2, it, lt, vm = engine.search_batch(boards, turns)
    t0 = pc(); ev = evaluate(lt, it, td, tp1, tp2); t_eval += pc() - t0
    engine.backprop_batch(*ev, it)
    remaining = n_playout - 1
    while remaining > 0:
        cur = min(K, remaining)
        remaining -= cur
        lb, td, tp1, tp2, it, lt, sym, vm = engine.search_batch_vl(cur, boards, turns)
        t0 = pc(); ev = evaluate(lt, it, td, tp1, tp2); t_eval += pc() - t0
        engine.backprop_batch_vl(cur, *ev, it, sym)
    return t_eval               # seconds spent in the numpy stand-in evaluator (not in the engine's entry points)


def step_io_bytes(n, n_playout, K, S, A):
    """(h2d, d2h) bytes per step moved by the host API."""
    h2d = d2h = 0
    ks = [1]
    rem = n_playout - 1
    while rem > 0:
        c = min(K, rem); rem -= c; ks.append(c)
    for i, k in enumerate(ks):
        rows = n * k
        h2d += n * S + n * 4                                   # boards + turns in
        d2h += rows * (S + 12 + 1 + 4 + A) + (rows * 4 if i else 0)   # leaves (+ sym ids for VL)
        h2d += rows * (A * 4 + 16 + 1) + (rows * 4 if i else 0)       # policy, wdl, ml, is_term (+ sym ids)
    h2d += n * 4                                               # prune/reset actions
    return h2d, d2h


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(len(r) > 2 + j and r[2 + j] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference engine on the host cores
# ------------------------------------------------------------------------------------------------------------
def run_reference(n, n_playout, K, steps, warmup, seed=0):
    """Times the reference's own CPU implementation (oracle/_ref/timing, -O3, OpenMP over all host threads; falls back
    to the single-thread C restatement when the build did not travel) on the same workload."""
    import oracle
    boards, turns = c4_random_roots(n, seed)
    A = 7
    if oracle.ref_available("timing"):
        mcts_cpp, _ = oracle.load_ref("timing")
        eng = mcts_cpp.BatchedMCTS_Connect4(n)
        kind, cores = "reference", os.cpu_count()
    else:
        eng = oracle.OracleMCTS("Connect4", n)
        kind, cores = "port", 1
    cfg = eng.config
    for k, v in SERVER_DEFAULTS.items():
        setattr(cfg, k, v)
    eng.set_seed(0)
    reset = np.full(n, -1, np.int32)
    for _ in range(warmup):
        host_step(eng, boards, turns, n_playout, K, A, reset)
    t0, t_ev = time.perf_counter(), 0.0
    for _ in range(steps):
        t_ev += host_step(eng, boards, turns, n_playout, K, A, reset)
    dt = time.perf_counter() - t0
    # The numpy stand-in evaluator between search and back-prop is not the reference engine's work (this repo's arm evaluates on
    # the device): `value` counts the time inside the reference's own entry points only; the whole loop is reported beside it.
    out = dict(value=n * n_playout * steps / max(dt - t_ev, 1e-9), ms_per_step=1e3 * (dt - t_ev) / steps, kind=kind, cores=cores,
               value_with_numpy_evaluator=n * n_playout * steps / dt, ms_per_step_with_numpy_evaluator=1e3 * dt / steps)
    native = run_reference_native(boards, turns, n, n_playout, K, steps, warmup)
    if native:
        out["native_harness"] = native
    return out


def run_reference_native(boards, turns, n, n_playout, K, steps, warmup):
    """The same loop over the reference's BatchedMCTS.h with no Python in between (oracle/_ref/timing/ref_native_bench, built from
    oracle/ref_native_bench.cpp against the reference headers).  Reported beside the pybind figure; None when the
    binary did not travel."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "timing", "ref_native_bench")
    if not os.path.exists(exe):
        return None
    try:
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(np.ascontiguousarray(boards, np.int8).tobytes())
            f.write(np.ascontiguousarray(turns, np.int32).tobytes())
            f.flush()
            env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))
            r = subprocess.run([exe, f.name, str(n), str(n_playout), str(K), str(steps), str(warmup)], capture_output=True, text=True,
                               env=env, timeout=600, check=True)
        j = json.loads(r.stdout)
        sims = n * n_playout * steps
        return dict(value=sims / j["engine_s"], value_whole_loop=sims / j["total_s"], threads=j["threads"],
                    note="reference BatchedMCTS<Connect4> driven from C++ (no pybind / numpy); value = time inside its entry points")
    except Exception as e:   # pragma: no cover
        return dict(error=repr(e)[:200])


def kernel_src_sha():
    """sha256 over the kernel sources: keys the committed ncu DRAM-traffic summary to the code it was captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "alphazero-al_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def load_traffic():
    """profiles/dram_traffic.json: {"kernel_src_sha", "captured_from", "kernels": {name: {"dram_read_bytes", "dram_write_bytes", "launches"}}}
    written by tools/ncu_summary.py from an `ncu --set full` capture of this very command.  Stale (other kernel sources) -> flagged."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    except Exception:
        return None
    t["stale"] = t.get("kernel_src_sha") != kernel_src_sha()
    return t


def reference_actor(timeout=900):
    """BASELINE config 1 on this box with the reference's own actor: unmodified Game.batch_self_play + AlphaZeroPlayer + its CNN's GPU
    predict (bf16 autocast) on its compiled engine (oracle/_ref: pysrc + timing build) - and the SAME unmodified actor code on this
    repository's modules (the four shims of INTEGRATION.md: device-resident search, the reference CNN evaluated through
    ReferenceNetAdapter with no host copy).  None when the byte-compiled layer is absent."""
    try:
        from oracle import refstack
        if not refstack.available("timing"):
            return None
        import tempfile
        res = {}
        with tempfile.TemporaryDirectory() as d:
            for key, engine in (("reference", "reference"), ("drop_in", "ours+wrapper")):
                ov = refstack.make_overlay(os.path.join(d, key), engine, "timing")
                out = os.path.join(d, key + ".json")
                env = {"OMP_NUM_THREADS": str(os.cpu_count())}
                refstack.run_driver(ov, "actor", out, dict(game="Connect4", n_games=100, n_playout=200, K=4, reps=1, warm_games=100), timeout=timeout,
                                    env_extra=env)
                res[key] = json.load(open(out))
        res["what"] = ("BASELINE config 1 through the reference's unmodified actor code: Game.batch_self_play(100 games, n=200, vl_batch=4, temp 1 "
                       "for 20 plies, td_steps 10), random-init reference CNN on the GPU.  reference = its compiled engine on all host cores + its "
                       "predict(); drop_in = the same code on this repository's modules (src/mcts_cpp, src/env_cpp, src/MCTS_cpp shims)")
        if res["reference"].get("games_per_sec") and res["drop_in"].get("games_per_sec"):
            res["drop_in_over_reference"] = res["drop_in"]["games_per_sec"] / res["reference"]["games_per_sec"]
        return res
    except Exception as e:   # pragma: no cover
        return {"error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 300 = about 1.2 s; reference arm: 5)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games-per-gpu", type=int, default=65536)
    ap.add_argument("--n-playout", type=int, default=200)
    ap.add_argument("--vl-batch", type=int, default=4)
    ap.add_argument("--cpu-baseline-games", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-actor-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--no-cnn", action="store_true")
    ap.add_argument("--cnn-slots", type=int, default=4096)
    ap.add_argument("--cnn-plies", type=int, default=3)
    ap.add_argument("--cnn-plies-100", type=int, default=30)
    ap.add_argument("--cnn-cached-plies", type=int, default=8)
    ap.add_argument("--cnn-cached-warm-plies", type=int, default=22)
    ap.add_argument("--selfplay-slots", type=int, default=65536)
    ap.add_argument("--selfplay-plies", type=int, default=40)
    ap.add_argument("--exchange-every", type=int, default=10, help="self-play plies per trajectory all-gather")
    ap.add_argument("--strong-total", type=int, default=65536, help="games in total for the strong-scaling self-play leg (BASELINE config 5)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per tree (Connect4: 1/2/4/8, 0 = auto)")
    ap.add_argument("--shards", type=int, default=0, help="independent tree shards on their own streams (0 = auto)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 300 if args.impl == "b200" else 5
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    G, n_playout, K = args.games_per_gpu, args.n_playout, args.vl_batch
    workload = f"connect4_mcts_n{n_playout}_k{K}_fresh_midgame_roots_{G}_games_per_gpu_constant_evaluator"
    config = {"workload": workload, "game": "Connect4", "games_per_gpu": G, "n_playout": n_playout, "vl_batch": K,
              "search_params": "server defaults (c_init 1.4, c_base 1000, fpu 0.2, alpha 0.3, eps 0.25, mlh 0.1/0.2, symmetry on)",
              "evaluator": "constant (uniform prior, fixed WDL/aux; stands in for the random-init CNN)",
              "parallelism": "independent game shards per GPU, no collective in the search path (trajectory all-gather in the self-play leg)",
              "l2": "inputs larger than L2: the tree arenas touched per step exceed the 126 MB L2, no flush between steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun exports OMP_NUM_THREADS=1; the reference engine is OpenMP-parallel and must get every host core
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
        r = run_reference(G, n_playout, K, args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": f"{G} games x {n_playout} sims x {args.steps} steps (full per-GPU workload)"},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "value_with_numpy_evaluator": r["value_with_numpy_evaluator"],
                "ms_per_step_with_numpy_evaluator": r["ms_per_step_with_numpy_evaluator"],
                "native_harness": r.get("native_harness"),
                "note": "value / ms_per_step count the time inside the reference engine's own entry points (search_batch[_vl], "
                        "backprop_batch[_vl], prune_roots: all host threads); the numpy stand-in evaluator that runs between them on "
                        "one thread is excluded (this repo's arm evaluates on the device) and reported in the *_with_numpy_evaluator keys"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # NCCL prints its version banner to STDOUT: keep stdout to the one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
    ds = importlib.import_module("alphazero-al_b200.device_search")

    A, S = 7, 42
    boards_np, turns_np = c4_random_roots(G, seed=1000 + rank)
    boards = torch.from_numpy(boards_np).to(dev)
    turns = torch.from_numpy(turns_np).to(dev)
    reset_np = np.full(G, -1, np.int32)
    eng = mcts_cpp.BatchedMCTS_Connect4(G, device=local_rank)
    for k, v in SERVER_DEFAULTS.items():
        setattr(eng.config, k, v)
    eng.set_seed(rank)
    if args.lanes:
        eng.set_lanes(args.lanes)
    buf = ds.LeafBuffers(G, G * K, A, (6, 7), dev)
    ev = ds.SyntheticEvaluator("Connect4", "constant")
    stream = torch.cuda.current_stream().cuda_stream

    shards = args.shards if args.shards > 0 else ds.auto_shards(G)

    def dev_step(n_shards=None):
        eng.reset_all_dev(stream)        # every tree back to a fresh root: stream-ordered, host arena bookkeeping included (no sync)
        buf.pack_roots(boards, turns, stream)
        return 2 + ds.playout_device(eng, buf, n_playout, K, ev, stream, shards=n_shards or shards)

    # ---- untimed pass with counters on: tree statistics for the roofline model ----
    eng.enable_stats(True)
    dev_step()
    torch.cuda.synchronize()
    st = eng.get_stats()
    eng.enable_stats(False)
    sims = max(st["sims"], 1)
    d_bar, E_bar, b_bar, x_bar = st["depth"] / sims, st["edges_scanned"] / sims, st["edges_created"] / sims, st["expansions"] / sims
    bytes_select_sim = 36 * d_bar + 40 * E_bar + 8 * (d_bar + 1) + 68 * x_bar + (S + A + 21)
    bytes_total_sim = 36 * d_bar + 40 * E_bar + 60 * (d_bar + 1) + 16 * b_bar + S + 5 * A + 118     # SURVEY.md 8(d) B_sim
    bytes_backprop_sim = bytes_total_sim - bytes_select_sim

    for _ in range(args.warmup):
        dev_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    launches = 0
    for _ in range(args.steps):
        launches += dev_step()           # native loop, tree shards on their own streams, replayed from a CUDA graph
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t_start.elapsed_time(t_end)
    clk = clocks.stop() if rank == 0 else None
    # the tree kernels with CUDA events around every select and back-prop launch (az_mcts_time_select; the loop is then issued
    # launch by launch instead of replayed from its graph), timed ALONE: two more steps with the whole batch per launch on one
    # stream, so a launch's elapsed time is the kernel's own (in the sharded step it shares the SMs with other shards' kernels)
    eng.time_select(True)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(2):
        dev_step(n_shards=1)
    a1.record()
    torch.cuda.synchronize()
    alone_step_ms = a0.elapsed_time(a1) / 2
    sel_ms, sel_launches, sel_rows = eng.get_select_time()
    bp_ms, bp_launches, bp_rows = eng.get_backprop_time()
    eng.time_select(False)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    total_sims = world * G * n_playout * args.steps
    value = total_sims / (ms * 1e-3)

    # ---- e2e: the public wrapper API with HOST buffers (numpy boards in, visit counts out) ----
    # BatchedMCTS.batch_playout / get_visits_count are the calls src/player.py makes (src/player.py:333-343); every step
    # copies that step's boards/turns/actions host->device and reads the visit counts back.
    e2e = e2e_split = None
    if not args.no_e2e:
        bm = importlib.import_module("alphazero-al_b200.batched_mcts")
        wrap = bm.BatchedMCTS(G, SERVER_DEFAULTS["c_init"], SERVER_DEFAULTS["c_base"], SERVER_DEFAULTS["dirichlet_alpha"], n_playout,
                              game_name="Connect4", noise_epsilon=SERVER_DEFAULTS["noise_epsilon"],
                              fpu_reduction=SERVER_DEFAULTS["fpu_reduction"], use_symmetry=True, mlh_slope=SERVER_DEFAULTS["mlh_slope"],
                              mlh_cap=SERVER_DEFAULTS["mlh_cap"], device=local_rank)
        wrap.seed(rank)
        if args.lanes:
            wrap.mcts.set_lanes(args.lanes)

        def wrap_step():
            wrap.prune_roots(reset_np)
            wrap.batch_playout(ev, boards_np, turns_np, vl_batch=K)
            return wrap.get_visits_count()

        for _ in range(2):
            c = wrap_step()
        assert int(c.sum()) == G * (n_playout - 1), "every fresh root must hold n-1 child visits"
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e2e_steps = min(args.steps, 60)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wrap_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": world * G * n_playout * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": G * (S + 4 + 4), "d2h_bytes_per_step": G * A * 8,
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "api": "BatchedMCTS.prune_roots + batch_playout(numpy boards, numpy turns) + get_visits_count() (wrapper mirror of "
                      "src/MCTS_cpp.py; evaluator runs on the device; az_mcts_playout_synthetic_host: boards staged / copied / searched / "
                      "counted shard by shard on 8 streams, int64 counts copied straight into the returned array's pinned memory)"}
        del wrap
        # secondary: the split host-buffer plugin API (search_batch[_vl] / backprop_batch[_vl] with numpy leaves every iteration)
        if world == 1 and not args.no_split:
            Gs = min(G, 8192)
            engs = mcts_cpp.BatchedMCTS_Connect4(Gs, device=local_rank)
            for k, v in SERVER_DEFAULTS.items():
                setattr(engs.config, k, v)
            bs_np, ts_np, rs_np = boards_np[:Gs], turns_np[:Gs], reset_np[:Gs]
            host_step(engs, bs_np, ts_np, n_playout, K, A, rs_np)
            t0, t_ev = time.perf_counter(), 0.0
            for _ in range(3):
                t_ev += host_step(engs, bs_np, ts_np, n_playout, K, A, rs_np)
            dts, t_ev = (time.perf_counter() - t0) / 3, t_ev / 3
            h2d, d2h = step_io_bytes(Gs, n_playout, K, S, A)
            e2e_split = {"value": Gs * n_playout / dts, "unit": UNIT, "games": Gs, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": 1e3 * dts, "ms_in_engine_calls": 1e3 * (dts - t_ev), "ms_in_numpy_evaluator": 1e3 * t_ev,
                         "value_engine_calls_only": Gs * n_playout / max(dts - t_ev, 1e-9),
                         "api": "mcts_cpp.search_batch[_vl]/backprop_batch[_vl] with numpy leaf buffers + numpy evaluator on the host "
                                "(what the unmodified src/MCTS_cpp.py drives)"}
            del engs
    del eng, buf
    torch.cuda.empty_cache()

    # ---- self-play games/s: the on-device driver (search + sample + record + env step + re-root + finished trajectories), with the
    # trajectory all-gather - the one collective of the path - inside the timed loop, issued on a side stream behind the next batch ----
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")

    def selfplay_leg(n_slots, plies, every):
        lo, _ = sp_mod.shard_range(world * n_slots, rank, world)
        sp = sp_mod.SelfPlay("Connect4", n_slots, n_playout, K, ds.SyntheticEvaluator("Connect4", "constant"), search_cfg=SERVER_DEFAULTS,
                             temperature=1.0, temp_decay_moves=20, temp_endgame=0.0, td_steps=10, seed=0, uid_base=lo,
                             uid_stride=world * n_slots, device=local_rank, out_capacity=2 * n_slots)
        sp.engine.reserve(16384)                  # arena compaction at re-roots keeps every game inside 16384 slots (2 pools x 512 KB)
        exch = sp_mod.TrajectoryExchange("Connect4", sp.out_capacity, dev)
        for _ in range(12):                       # reach the steady state of continuously restarting games
            sp.ply()
        exch.gather(sp.hand_over())               # warm-up exchange: NCCL connection set-up, buffer registration
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        p0, l0 = sp.plies, sp.launches
        t0 = time.perf_counter()
        pending, gathered_games, gathered_pos, replay_rows = None, 0, 0, 0

        def consume():                            # the receiving side: oldest finished gather -> training tuples, on the exchange's stream
            nonlocal gathered_games, gathered_pos, replay_rows
            rec = exch.collect()
            gathered_games += len(rec)
            gathered_pos += rec.positions
            if len(rec):
                with torch.cuda.stream(exch.stream):
                    replay_rows += int(rec.to_replay_tensors(10, stream=exch.stream.cuda_stream)["state"].shape[0])

        done_plies = 0
        while done_plies < plies:
            for _ in range(min(every, plies - done_plies)):
                sp.ply()
            done_plies += min(every, plies - done_plies)
            ring = sp.hand_over()                 # batch closed on the main stream; the other ring takes over
            if pending is not None:               # the previous batch's gather is issued behind this batch's (already enqueued) search
                exch.submit(pending)
                while len(exch.pending) > 1:
                    consume()
            pending = ring
        e_main = torch.cuda.Event()
        e_main.record()
        exch.submit(pending)
        e_main.synchronize()
        t_main = time.perf_counter() - t0
        while exch.pending:
            consume()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tot = torch.tensor([dt, t_main], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dt, t_main = float(tot[0].item()), float(tot[1].item())
        out = {"games_per_sec": gathered_games / dt, "sims_per_sec": world * n_slots * n_playout * (sp.plies - p0) / dt,
               "slots_per_gpu": n_slots, "plies_timed": sp.plies - p0, "games_finished_all_ranks": gathered_games,
               "positions_gathered": gathered_pos, "replay_rows_expanded": replay_rows, "seconds": dt,
               "exchange_every_plies": every, "exchanges": len(exch.ms) - 1,
               "allgather_ms_mean": float(np.mean(exch.ms[1:])) if len(exch.ms) > 1 else None,
               "allgather_ms_max": float(np.max(exch.ms[1:])) if len(exch.ms) > 1 else None,
               "collective_exposed_ms": 1e3 * (dt - t_main),
               "bytes_received_per_rank": exch.bytes_gathered, "position_record_bytes": sp.pb, "game_header_bytes": sp_mod.GAME_BYTES,
               "gpu_launches": sp.launches - l0}
        del sp, exch
        torch.cuda.empty_cache()
        return out

    selfplay = None
    if not args.no_selfplay:
        selfplay = selfplay_leg(min(G, args.selfplay_slots), args.selfplay_plies, args.exchange_every)
        selfplay["note"] = ("continuous self-play with tree reuse, temp 1 for 20 plies then 0, constant evaluator; games counted from the "
                            "gathered records of all ranks (finished games, no estimate); every `exchange_every_plies` plies the finished "
                            "trajectories (compact records: 32-byte header + 64 bytes per position) are all-gathered over NCCL on a side "
                            "stream while the next batch is searched, then expanded into replay tensors on every rank; "
                            "collective_exposed_ms = time after the last ply until the last gather + expansion completed")
        if world > 1 and args.strong_total >= world * 1024:
            selfplay["strong_scaling"] = dict(selfplay_leg(args.strong_total // world, args.selfplay_plies, args.exchange_every),
                                              total_games_in_flight=args.strong_total,
                                              note="BASELINE config 5 as stated: 65 536 concurrent games in total, split over the GPUs")
    # ---- the same self-play with a random-init CNN of the reference's Connect4 architecture in the loop (bf16 autocast) ----
    selfplay_cnn = None
    if not args.no_cnn and world == 1:
        nets = importlib.import_module("alphazero-al_b200.nets")
        torch.manual_seed(0)
        net = nets.C4Net(device=f"cuda:{local_rank}")
        n_slots = args.cnn_slots
        sp = sp_mod.SelfPlay("Connect4", n_slots, n_playout, K, net, search_cfg=SERVER_DEFAULTS, temperature=1.0, temp_decay_moves=20,
                             temp_endgame=0.0, td_steps=10, seed=0, device=local_rank, out_capacity=4 * n_slots)
        sp.engine.reserve(16384)
        sp.ply()
        torch.cuda.synchronize()
        p0 = sp.plies
        t0 = time.perf_counter()
        for _ in range(args.cnn_plies):
            sp.ply()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        selfplay_cnn = {"sims_per_sec": n_slots * n_playout * (sp.plies - p0) / dt, "positions_per_sec": n_slots * (sp.plies - p0) / dt,
                        "slots": n_slots, "plies_timed": sp.plies - p0,
                        "evaluator": "C4Net (160358 params, reference Connect4 CNN shape), random init, bf16 autocast, device contract "
                                     "(leaves -> planes -> net -> finalize, no host copy)"}
        del sp
        # BASELINE config 1 itself: 100 games, n=200, K=4, random-init CNN - 100..400 leaves per evaluation, bound by the launches
        # of the forward pass; NetEvaluator replays it from a CUDA graph (eager timed beside it).  games/s = finished games / time
        # over the timed plies at the steady state of continuously restarting slots.
        c0 = {}
        for name, graph_rows, csize in (("eager", 0, 0), ("graph", 8192, 0), ("graph+cache", 8192, 1 << 20)):
            sp = sp_mod.SelfPlay("Connect4", 100, n_playout, K, net, search_cfg=SERVER_DEFAULTS, temperature=1.0, temp_decay_moves=20,
                                 temp_endgame=0.0, td_steps=10, seed=0, device=local_rank, out_capacity=1024, cache_size=csize)
            sp.evaluator.graph_rows = graph_rows
            for _ in range(14):                   # past the opening: slots at mixed plies (and the cache past the shared first moves)
                sp.ply()
            sp.drain()
            torch.cuda.synchronize()
            p0, t0 = sp.plies, time.perf_counter()
            for _ in range(args.cnn_plies_100 if name != "eager" else max(4, args.cnn_plies_100 // 4)):
                sp.ply()
            fin = sp.drain()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            c0[name] = {"sims_per_sec": 100 * n_playout * (sp.plies - p0) / dt, "ms_per_ply": 1e3 * dt / (sp.plies - p0),
                        "games_finished": len(fin), "games_per_sec": len(fin) / dt, "graph_replays": sp.evaluator.graph_replays}
            del sp
        selfplay_cnn["config1_100_games"] = c0
        # the same with the device evaluation cache + in-batch de-duplication: only distinct unseen positions
        # reach the network.  Timed at steady state (slots at mixed plies), not on the opening, where nearly every leaf repeats.
        if args.cnn_cached_plies > 0:
            sp = sp_mod.SelfPlay("Connect4", n_slots, n_playout, K, net, search_cfg=SERVER_DEFAULTS, temperature=1.0, temp_decay_moves=20,
                                 temp_endgame=0.0, td_steps=10, seed=0, device=local_rank, out_capacity=16 * n_slots, cache_size=1 << 22)
            sp.engine.reserve(16384)
            for _ in range(args.cnn_cached_warm_plies):
                sp.ply()
            sp.drain()
            torch.cuda.synchronize()
            st0, r0, p0 = sp.eval_cache.stats(), sp.evaluator.net_rows, sp.plies
            t0 = time.perf_counter()
            for _ in range(args.cnn_cached_plies):
                sp.ply()
            fin = sp.drain()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st1 = sp.eval_cache.stats()
            look = max(st1["lookups"] - st0["lookups"], 1)
            selfplay_cnn["cached"] = {
                "sims_per_sec": n_slots * n_playout * (sp.plies - p0) / dt, "positions_per_sec": n_slots * (sp.plies - p0) / dt,
                "games_finished": len(fin), "games_per_sec": len(fin) / dt, "plies_timed": sp.plies - p0, "plies_before": p0,
                "cache_entries": st1["capacity"], "hit_frac": (st1["hits"] - st0["hits"]) / look, "dup_frac": (st1["dups"] - st0["dups"]) / look,
                "net_rows_frac": (sp.evaluator.net_rows - r0) / look,
                "note": "device evaluation cache + in-batch de-duplication; timed after the warm plies, slots at mixed game plies"}
            del sp
        del net

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic_all = load_traffic()

    def kernel_roofline(kname, key, ms_sum, n_launch, rows, bytes_sim):
        ach = (rows * bytes_sim) / (ms_sum * 1e-3) / 1e9 if ms_sum > 0 else 0.0
        tr = src = None
        if traffic_all and not traffic_all["stale"] and key in traffic_all.get("kernels", {}):
            k = traffic_all["kernels"][key]
            tr = (k["dram_read_bytes"] + k["dram_write_bytes"]) / max(k["launches"], 1)
            src = f"profiles/dram_traffic.json ({traffic_all.get('captured_from')}; kernel_src_sha {traffic_all.get('kernel_src_sha')})"
        elif traffic_all and traffic_all["stale"]:
            src = "profiles/dram_traffic.json is stale (captured from other kernel sources): not reported"
        return {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": tr,
                "traffic_source": src, "us_per_launch": 1e3 * ms_sum / max(n_launch, 1), "launches_timed": n_launch,
                "algorithmic_bytes_per_sim": bytes_sim, "algorithmic_bytes_per_launch": bytes_sim * rows / max(n_launch, 1),
                "share_of_one_stream_step": ms_sum / (2 * alone_step_ms) if alone_step_ms > 0 else None}

    lanes, variant = 1, 1
    sel = kernel_roofline("az::k_select_f<C4,VL,AUX,RO>", "k_select_f", sel_ms, sel_launches, sel_rows, bytes_select_sim)
    bp = kernel_roofline("az::k_backprop_f<C4,VL,RO>", "k_backprop_f", bp_ms, bp_launches, bp_rows, bytes_backprop_sim)
    dominant = sel if sel_ms >= bp_ms else bp
    roofline = dict(dominant)
    roofline.update({
        "kernels": {"select": sel, "backprop": bp},
        "note": ("per kernel: achieved = algorithmic bytes (SURVEY.md 8d model x on-device tree statistics of an untimed pass) / CUDA-event "
                 "time of its launches, events on the launching stream (az_mcts_time_select), taken in 2 extra steps right after the timed "
                 "region with the whole batch per launch on ONE stream so that a launch's elapsed time is the kernel's own (the timed region "
                 "replays the step from a CUDA graph with tree shards on several streams, where kernels of different shards share the SMs).  "
                 "The top-level keys repeat the kernel with the larger share of the step"),
        "ms_per_step_one_stream": alone_step_ms, "shards_in_timed_region": shards,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "bytes_per_sim_whole_path": bytes_total_sim,
        "tree_stats": {"depth": d_bar, "edges_scanned": E_bar, "edges_created": b_bar, "expansions": x_bar},
        "whole_path_frac": value / world * bytes_total_sim / 1e9 / peak})
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--games-per-gpu",
                                  str(args.cpu_baseline_games), "--n-playout", str(n_playout), "--vl-batch", str(K),
                                  "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                                 env={k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK")})
            r = json.loads(out.stdout.strip().splitlines()[-1])
            cpu = r["cpu_baseline"]
            cpu["sample"] = (f"{args.cpu_baseline_games} games x {n_playout} sims x 3 steps (same workload, fewer games); time inside the "
                             "reference engine's entry points only, the numpy stand-in evaluator excluded")
            cpu["value_with_numpy_evaluator"] = r.get("value_with_numpy_evaluator")
            cpu["native_harness"] = r.get("native_harness")
        except Exception as e:   # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        if not args.no_actor_baseline:
            cpu["reference_actor_config1"] = reference_actor()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "clocks": clk, "e2e": e2e, "e2e_split_api": e2e_split, "selfplay": selfplay, "selfplay_cnn": selfplay_cnn, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
