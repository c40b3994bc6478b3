"""The reference's own Python layer (src/MCTS_cpp.py, src/game.py, src/player.py, src/ReplayBuffer.py, src/environments/...)
run UNMODIFIED on either engine - TEST INFRASTRUCTURE ONLY (tests/, and bench.py's reference / cpu_baseline legs).

``make -C oracle ref`` byte-compiles that layer into ``oracle/_ref/pysrc/src`` (oracle/compile_pysrc.py; no source is copied, the
``.pyc`` files are git-ignored like the compiled ``.so`` modules and travel to the GPU box with them).  ``make_overlay`` builds a
directory whose ``src`` package is made of copies of those byte-code images (named ``<module>.pyc``) plus what the reference's build step would have put there:

* engine "reference": ``src/mcts_cpp*.so`` / ``src/env_cpp*.so`` -> oracle/_ref/<kind>/ (what setup.py:62-65 does), or
* engine "ours": the three shims of INTEGRATION.md section 1 (``src/azb200`` -> the package, ``src/mcts_cpp.py``, ``src/env_cpp/``);
* engine "ours+wrapper": those plus the fourth shim, ``src/MCTS_cpp.py`` -> ``batched_mcts.BatchedMCTS`` (the mirror of the search wrapper
  whose device-resident path keeps the whole per-move loop on the GPU), replacing the reference's own wrapper module.

``run_driver`` executes tests/refstack_driver.py in a subprocess with that overlay first on PYTHONPATH (two overlays cannot share a
process: both define the package ``src``)."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
PYSRC = os.path.join(_HERE, "_ref", "pysrc", "src")


def available(kind: str = "parity") -> bool:
    d = os.path.join(_HERE, "_ref", kind)
    return os.path.isfile(os.path.join(PYSRC, "MCTS_cpp.pycode")) and os.path.isdir(d) and any(f.startswith("mcts_cpp") for f in os.listdir(d))


def make_overlay(dst: str, engine: str, kind: str = "parity") -> str:
    """Creates <dst>/src and returns <dst> (to be put first on PYTHONPATH)."""
    src = os.path.join(dst, "src")
    os.makedirs(src, exist_ok=True)
    for root, dirs, files in os.walk(PYSRC):
        rel = os.path.relpath(root, PYSRC)
        os.makedirs(os.path.join(src, rel), exist_ok=True)
        for f in files:
            if f.endswith(".pycode") and f != "pipeline.pycode":       # byte-code images -> sourceless modules of the scratch overlay
                if engine == "ours+wrapper" and rel == "." and f == "MCTS_cpp.pycode":
                    continue                                        # replaced by the shim below
                shutil.copyfile(os.path.join(root, f), os.path.join(src, rel, f[:-7] + ".pyc"))
    if engine == "reference":
        d = os.path.join(_HERE, "_ref", kind)
        for f in os.listdir(d):
            if f.endswith(".so"):
                os.symlink(os.path.join(d, f), os.path.join(src, f))
    elif engine in ("ours", "ours+wrapper"):
        if engine == "ours+wrapper":
            with open(os.path.join(src, "MCTS_cpp.py"), "w") as f:
                f.write("from src.azb200.batched_mcts import BatchedMCTS, _default_convert_board, _relative_wdl_to_absolute\n")
        os.symlink(os.path.join(ROOT, "alphazero-al_b200"), os.path.join(src, "azb200"))
        with open(os.path.join(src, "mcts_cpp.py"), "w") as f:
            f.write("from src.azb200.mcts_cpp import *\n")
        os.makedirs(os.path.join(src, "env_cpp"))
        with open(os.path.join(src, "env_cpp", "__init__.py"), "w") as f:
            f.write("from src.azb200.env_cpp import *\n")
        for g in ("connect4", "othello", "gomoku"):
            with open(os.path.join(src, "env_cpp", f"{g}.py"), "w") as f:
                f.write(f"from src.azb200.env_cpp.{g} import *\n")
    else:
        raise ValueError(engine)
    return dst


def run_driver(overlay: str, task: str, out_path: str, params: dict, timeout: int = 1800, env_extra: dict | None = None) -> str:
    """python tests/refstack_driver.py <task> <out_path> <json params> with the overlay's ``src`` package.  Returns stdout."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([overlay, ROOT, os.path.join(ROOT, "tests")]), PYTHONDONTWRITEBYTECODE="1")
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refstack_driver.py"), task, out_path, json.dumps(params)],
                       cwd=overlay, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"refstack_driver {task} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}")
    return r.stdout
