"""CPU-only checks: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares; the
host-side mirror validates arguments like the reference bindings; there is no CPU fallback."""
import ctypes
import glob
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(az_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_builds_and_exports_every_declared_symbol():
    libmod = importlib.import_module("alphazero-al_b200._lib")
    path = libmod.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    lib.az_version.restype = ctypes.c_char_p
    assert lib.az_version().decode().startswith("azb200")


def test_sm100a_code_is_embedded():
    import subprocess
    libmod = importlib.import_module("alphazero-al_b200._lib")
    out = subprocess.run(["cuobjdump", "-lelf", libmod.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_static_traits_and_config_defaults_match_reference():
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    assert (m.BatchedMCTS_Connect4.action_size, m.BatchedMCTS_Connect4.board_size, m.BatchedMCTS_Connect4.board_shape) == (7, 42, (6, 7))
    assert (m.BatchedMCTS_Othello.action_size, m.BatchedMCTS_Othello.board_size, m.BatchedMCTS_Othello.board_shape) == (65, 64, (8, 8))
    c = m.SearchConfig()                                   # src/cpp/MCTSNode.h:47-61
    assert (c.c_init, c.c_base, c.noise_epsilon, c.mlh_slope, c.score_utility_factor, c.score_scale, c.value_decay) == \
           (1.25, 19652.0, 0.25, 0.0, 0.0, 8.0, 1.0)
    assert abs(c.dirichlet_alpha - 0.3) < 1e-7 and abs(c.fpu_reduction - 0.4) < 1e-7 and abs(c.mlh_cap - 0.2) < 1e-7
    assert c.use_symmetry is True and c.vl_count == 1
    c.use_symmetry = False
    c.vl_count = 3
    assert c.use_symmetry is False and c.vl_count == 3
    with pytest.raises(AttributeError):
        c.no_such_field = 1
    with pytest.raises(TypeError):
        m.IEvaluator_Othello()
    assert isinstance(m.RolloutEvaluator_Othello(), m.IEvaluator_Othello)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = importlib.import_module("alphazero-al_b200.mcts_cpp")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.BatchedMCTS_Connect4(8)
    env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env_cpp.BatchedEnv("Connect4", 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env_cpp.BatchedGomoku(8, 15, 5)


def test_hash_evaluator_is_flip_equivariant_and_deterministic():
    ev = importlib.import_module("alphazero-al_b200.evaluators")
    from harness import random_positions
    b, t = random_positions("Connect4", 64, 30, 5)
    e = ev.HashEvaluator("Connect4", "equivariant")
    p, w, a = e.raw(b, t)
    pf, wf, af = e.raw(b[:, :, ::-1].copy(), t)
    assert np.array_equal(p[:, ::-1], pf) and np.array_equal(w, wf) and np.array_equal(a, af)
    p2, _, _ = ev.HashEvaluator("Connect4", "hash").raw(b, t)
    assert np.array_equal(p2, ev.HashEvaluator("Connect4", "hash").raw(b, t)[0]) and not np.array_equal(p2, p)
