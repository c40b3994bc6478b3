"""bench.py contract, CPU side: the reference arm (`--impl reference`) prints one JSON line with the keys the driver reads, under
plain python and as rank 0 of a 2-rank launch (the other rank prints nothing), and the b200 arm refuses to run without a CUDA
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
ARGS = ["--impl", "reference", "--games-per-gpu", "256", "--n-playout", "40", "--steps", "1", "--warmup", "1"]


def _run(extra_env=None, args=ARGS):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcts_simulations_per_sec" and d["unit"] == "sims/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["kind"] == ("reference" if oracle.ref_available("timing") else "port")
    assert d["config"]["workload"].startswith("connect4_mcts_n40_k4") and "model" not in d["config"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "timing", "ref_native_bench")):
        nh = d["native_harness"]
        assert "error" not in nh and nh["value"] > 0 and nh["value_whole_loop"] > 0 and nh["threads"] >= 1


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = _run(args=["--steps", "1", "--warmup", "1", "--games-per-gpu", "64"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
