"""Deterministic synthetic leaf evaluators (host numpy versions).

They stand in for ``CNN.predict`` (src/environments/Connect4/Network.py:267-288 - same I/O contract:
symmetrised leaf boards + side to move in, policy over A actions + WDL + aux out) wherever a test or
benchmark needs bit-reproducible priors: SURVEY.md section 4 item 2 ("constant prior; hash-of-board prior;
symmetry-equivariant prior").  Each has a CUDA twin in ``csrc/az_eval.cu`` that returns identical bits, so the
device-resident search loop can be checked against the host-buffer loop.

All arithmetic that reaches the engine is exact in fp32 (small integers scaled by powers of two, one IEEE
division), hence identical on numpy / glibc / CUDA.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def _bit_positions(shape):
    """Engine bit index of row-major cell j (csrc/az_games.cuh cell_bit): Connect4 col*7+(5-row), Othello j."""
    R, Cc = shape
    if (R, Cc) == (6, 7):
        return np.array([c * 7 + (5 - r) for r in range(6) for c in range(7)], dtype=np.uint64)
    return np.arange(R * Cc, dtype=np.uint64)


def _bitboards(boards: np.ndarray):
    """(own1, own2) uint64 per board in the engine's bitboard layout (stones of +1 / -1)."""
    B = boards.shape[0]
    flat = boards.reshape(B, -1)
    w = (np.uint64(1) << _bit_positions(boards.shape[1:]))[None, :]
    own1 = ((flat == 1).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    own2 = ((flat == -1).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    return own1, own2


def board_key(own1, own2, turns):
    t = np.where(np.asarray(turns) == 1, np.uint64(0x5555555555555555), np.uint64(0xAAAAAAAAAAAAAAAA))
    with np.errstate(over="ignore"):
        return splitmix64(own1 ^ splitmix64((own2 + np.uint64(0x9E3779B97F4A7C15)) & _M64) ^ t)


class HashEvaluator:
    """hash-of-board prior / value.

    mode: "hash" (policy depends on the literal board), "equivariant" (Connect4 only: policy commutes with the
    horizontal flip, value/aux are flip-invariant, so the engine's random leaf symmetry cannot change visit
    counts) or "constant" (uniform policy, fixed value).
    Call signature mirrors what the search wrapper hands to backprop (src/MCTS_cpp.py:275-350):
        (leaf_boards int8[B,R,C], leaf_turns i32[B], is_term u8[B], term_d, term_p1w, term_p2w f32[B])
        -> (probs f32[B,A], d, p1w, p2w, moves_left f32[B])   terminal rows: policy 0, terminal WDL, ml 0.
    """

    def __init__(self, game: str = "Connect4", mode: str = "hash"):
        self.game, self.mode = game, mode
        self.A = 7 if game == "Connect4" else 65
        if mode == "equivariant" and game != "Connect4":
            raise ValueError("equivariant hash evaluator is defined for Connect4 only")

    def raw(self, boards, turns):
        """(probs[B,A], wdl_rel[B,3] = [draw, win(to move), loss(to move)], aux[B]) - the CNN.predict contract."""
        boards = np.asarray(boards, dtype=np.int8)
        turns = np.asarray(turns, dtype=np.int32)
        B, A = boards.shape[0], self.A
        if self.mode == "constant":
            probs = np.ones((B, A), np.float32)
            wdl = np.tile(np.array([0.25, 0.5, 0.25], np.float32), (B, 1))
            aux = np.full(B, 10.0 if self.game == "Connect4" else 0.125, np.float32)
            return probs, wdl, aux
        own1, own2 = _bitboards(boards)
        acts = np.arange(A, dtype=np.uint64)[None, :]
        if self.mode == "equivariant":
            f1, f2 = _bitboards(boards[:, :, ::-1])
            canon = (own1 < f1) | ((own1 == f1) & (own2 <= f2))      # literal board is the canonical one
            selfsym = (own1 == f1) & (own2 == f2)
            k1 = np.where(canon, own1, f1)
            k2 = np.where(canon, own2, f2)
            h = board_key(k1, k2, turns)
            a_can = np.where(canon[:, None], acts, np.uint64(A - 1) - acts)
            a_can = np.where(selfsym[:, None], np.minimum(a_can, np.uint64(A - 1) - a_can), a_can)
        else:
            h = board_key(own1, own2, turns)
            a_can = np.broadcast_to(acts, (B, A))
        with np.errstate(over="ignore"):
            ph = splitmix64((h[:, None] + a_can + np.uint64(1)) & _M64)
        probs = (((ph >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 65536.0)
        w = []
        for i, c in enumerate((0x1111, 0x2222, 0x3333)):
            wi = (splitmix64(h ^ np.uint64(c)) >> np.uint64(40)) & np.uint64(0xFF)
            w.append(wi.astype(np.float32) + np.float32(1.0))
        s = (w[0] + w[1]) + w[2]
        wdl = np.stack([w[0] / s, w[1] / s, w[2] / s], axis=1).astype(np.float32)
        if self.game == "Connect4":
            aux = ((h >> np.uint64(20)) & np.uint64(31)).astype(np.float32)
        else:
            aux = ((h >> np.uint64(20)) & np.uint64(63)).astype(np.float32) * np.float32(1.0 / 32.0) - np.float32(1.0)
        return probs.astype(np.float32), wdl, aux.astype(np.float32)

    def __call__(self, leaf_boards, leaf_turns, is_term, term_d, term_p1w, term_p2w):
        probs, wdl, aux = self.raw(leaf_boards, leaf_turns)
        t = np.asarray(is_term).astype(bool)
        p1 = np.asarray(leaf_turns) == 1
        d = np.where(t, term_d, wdl[:, 0]).astype(np.float32)
        p1w = np.where(t, term_p1w, np.where(p1, wdl[:, 1], wdl[:, 2])).astype(np.float32)
        p2w = np.where(t, term_p2w, np.where(p1, wdl[:, 2], wdl[:, 1])).astype(np.float32)
        ml = np.where(t, np.float32(0), aux).astype(np.float32)
        probs = np.where(t[:, None], np.float32(0), probs).astype(np.float32)
        return probs, d, p1w, p2w, ml
