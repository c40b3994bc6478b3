"""Compact trajectory records (include/azb200_selfplay.h) built on the host from a self-play fixture of the reference
(tests/golden/py_*_selfplay_*.npz, produced by the reference's own Game.batch_self_play).  Test helper."""
import importlib
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def fixture_records(z, game, uid0=0):
    """-> (Records on the CPU, uid order = fixture order).  bit layout: alphazero-al_b200/evaluators._bit_positions."""
    sp = importlib.import_module("alphazero-al_b200.selfplay")
    ev = importlib.import_module("alphazero-al_b200.evaluators")
    pb = sp.pos_bytes(game)
    state = z["state"]                                   # [P, 3, R, C] relative planes
    P = state.shape[0]
    A = z["prob"].shape[1]
    turn = state[:, 2, 0, 0].astype(np.int8)
    w = (np.uint64(1) << ev._bit_positions(state.shape[2:]))[None, :]
    own = (state[:, 0].reshape(P, -1).astype(np.uint64) * w).sum(1, dtype=np.uint64)
    opp = (state[:, 1].reshape(P, -1).astype(np.uint64) * w).sum(1, dtype=np.uint64)
    bb0 = np.where(turn == 1, own, opp)
    bb1 = np.where(turn == 1, opp, own)
    pos = np.zeros((P, pb), np.uint8)
    pos[:, 0:8] = bb0.view(np.uint8).reshape(P, 8)
    pos[:, 8:16] = bb1.view(np.uint8).reshape(P, 8)
    pos[:, 16:28] = np.ascontiguousarray(z["root_wdl"], np.float32).view(np.uint8).reshape(P, 12)
    pos[:, 28] = turn.view(np.uint8)
    if game == "Othello":                                # consecutive passes before each position (the action stream tells)
        act = z["action"]
        passes = np.zeros(P, np.uint8)
        start = 0
        for L in z["length"]:
            run = 0
            for t in range(L):
                passes[start + t] = run
                run = run + 1 if act[start + t] == 64 else 0
            start += L
        pos[:, 29] = passes
    pos[:, 32:32 + 4 * A] = np.ascontiguousarray(z["prob"], np.float32).view(np.uint8).reshape(P, 4 * A)
    m = len(z["length"])
    games = np.zeros((m, 4), np.int64)
    games[:, 0] = uid0 + np.arange(m)
    games[:, 1] = np.cumsum(z["length"]) - z["length"]
    g32 = games.view(np.int32).reshape(m, 8)
    g32[:, 4] = z["length"]
    g32[:, 5] = z["winner"]
    return sp.Records(game, torch.from_numpy(games.view(np.uint8).reshape(m, 32).copy()), torch.from_numpy(pos))


def forced_actions(z, game):
    """int8[games, max_plies] opening script that replays the fixture's games (-1 beyond a game's end)."""
    sp = importlib.import_module("alphazero-al_b200.selfplay")
    T = sp.max_plies(game)
    m = len(z["length"])
    out = np.full((m, T), -1, np.int8)
    start = 0
    for i, L in enumerate(z["length"]):
        out[i, :L - 1] = z["action"][start:start + L - 1]
        start += L
    return out
