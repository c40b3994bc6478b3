"""Generates the Python-layer fixtures from the UNMODIFIED reference (its Python layer byte-compiled into oracle/_ref/pysrc, its
engine and envs compiled into oracle/_ref/parity; both by `make -C oracle ref`).  Build container only:

    python tests/golden/make_golden_py.py

Everything is produced by the reference's own code (tests/refstack_driver.py only calls it): `Game.batch_self_play` +
`AlphaZeroPlayer.get_batch_action` training tuples, `ReplayBuffer.save` files, the actor's upload pickle, and
`BatchedMCTS.batch_playout` with the LRU cache on.  Deterministic: no Dirichlet noise, hash evaluator; leaf symmetry either off or on with
the flip-equivariant evaluator (then the reference's mt19937 symmetry ids cannot change a visit count).  Self-play samples the first
plies with temperature 1 from numpy's seeded global generator (src/player.py:364-369) so that the games of a batch differ; the
reference's Python layer on another engine consumes that generator identically as long as the visit counts are identical, and the
on-device driver is checked by replaying the recorded moves."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refstack  # noqa: E402

C4_CFG = dict(c_init=1.4, c_base=1000, fpu_reduction=0.2, mlh_slope=0.1, mlh_cap=0.2)
OTH_CFG = dict(c_init=1.4, c_base=500, fpu_reduction=0.2, score_utility_factor=0.15, score_scale=8.0)

SELFPLAY = {
    "py_c4_selfplay_k4_sym": dict(game="Connect4", n_games=24, n_playout=48, K=4, td_steps=3, mode="equivariant", use_symmetry=True,
                                  cfg=C4_CFG, seed=5, formats=True, temperature=1.0, temp_decay_moves=6),
    "py_c4_selfplay_k1_td0": dict(game="Connect4", n_games=12, n_playout=24, K=1, td_steps=0, mode="hash", use_symmetry=False,
                                  cfg=dict(C4_CFG, value_decay=0.98), seed=6, temperature=1.0, temp_decay_moves=4),
    "py_oth_selfplay_k4": dict(game="Othello", n_games=6, n_playout=24, K=4, td_steps=5, mode="hash", use_symmetry=False, cfg=OTH_CFG,
                               seed=7, formats=True, temperature=1.0, temp_decay_moves=10),
}
PLAYOUT = {
    "py_c4_playout_cache": dict(game="Connect4", n=16, n_playout=60, K=4, moves=3, cache_size=96, mode="hash", use_symmetry=False,
                                cfg=C4_CFG, seed=8, max_plies=10, pos_seed=21),
    "py_c4_playout_cache_sym_k1": dict(game="Connect4", n=8, n_playout=30, K=1, moves=2, cache_size=40, mode="equivariant",
                                       use_symmetry=True, cfg=C4_CFG, seed=9, max_plies=6, pos_seed=22),
    "py_oth_playout_cache": dict(game="Othello", n=6, n_playout=40, K=4, moves=2, cache_size=64, mode="hash", use_symmetry=False,
                                 cfg=OTH_CFG, seed=10, max_plies=20, pos_seed=23),
}


def main():
    assert refstack.available("parity"), "run `make -C oracle ref` first"
    with tempfile.TemporaryDirectory() as d:
        ov = refstack.make_overlay(os.path.join(d, "ref"), "reference", "parity")
        for name, p in SELFPLAY.items():
            print(refstack.run_driver(ov, "selfplay", os.path.join(HERE, name + ".npz"), p).strip())
        for name, p in PLAYOUT.items():
            print(refstack.run_driver(ov, "playout", os.path.join(HERE, name + ".npz"), p).strip())
    with open(os.path.join(HERE, "py_cases.json"), "w") as f:
        json.dump({"selfplay": SELFPLAY, "playout": PLAYOUT}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
