"""Pin the C restatement (oracle/az_oracle.c) against the UNMODIFIED reference engine compiled into
oracle/_ref/parity (-O2 -ffp-contract=off).  Everything deterministic must be bit-exact."""
import importlib

import numpy as np
import pytest

import oracle
from harness import SERVER_DEFAULTS, compare_engines, counts, playout, random_positions, set_config

ev_mod = importlib.import_module("alphazero-al_b200.evaluators")

pytestmark = pytest.mark.skipif(not oracle.ref_available("parity"), reason="oracle/_ref/parity not built")


def _pair(game, n):
    mcts_cpp, _ = oracle.load_ref("parity")
    ref = getattr(mcts_cpp, f"BatchedMCTS_{game}")(n)
    orc = oracle.OracleMCTS(game, n)
    return ref, orc


def _compare(game, n, n_playout, K, cfg, mode="hash", boards=None, turns=None, moves=1, compare_leaves=True):
    ref, orc = _pair(game, n)
    return compare_engines(ref, orc, game, n, n_playout, K, cfg, mode, boards, turns, moves, compare_leaves)


@pytest.mark.parametrize("K", [1, 2, 4, 8])
def test_c4_fresh_roots(K):
    c = _compare("Connect4", 24, 60, K, SERVER_DEFAULTS)
    assert (c.sum(axis=1) == 59).all()


def test_c4_remainder_and_tree_reuse():
    cfg = dict(SERVER_DEFAULTS, value_decay=0.97)
    boards, turns = random_positions("Connect4", 32, 20, 1)
    _compare("Connect4", 32, 51, 4, cfg, boards=boards, turns=turns, moves=12)


def test_c4_mlh_off_default_config():
    _compare("Connect4", 16, 80, 8, dict(dirichlet_alpha=0.0, use_symmetry=False), moves=3)


def test_c4_constant_evaluator_symmetric_counts():
    c = _compare("Connect4", 4, 200, 4, dict(dirichlet_alpha=0.0, use_symmetry=False), mode="constant")
    assert (c.sum(axis=1) == 199).all()


def test_c4_symmetry_on_equivariant_prior():
    # sym ids come from different RNG streams, but an equivariant evaluator makes the counts independent of them
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True)
    boards, turns = random_positions("Connect4", 32, 16, 2)
    _compare("Connect4", 32, 64, 4, cfg, mode="equivariant", boards=boards, turns=turns, moves=4, compare_leaves=False)


def test_c4_fresh_root_turn_minus_one_quirk():
    # fresh trees always get root.turn=+1 (MCTS.h:77-82) even when O is to move
    boards, turns = random_positions("Connect4", 16, 9, 3)
    _compare("Connect4", 16, 40, 4, SERVER_DEFAULTS, boards=boards, turns=turns, moves=1)


def test_c4_near_terminal_roots():
    boards, turns = random_positions("Connect4", 48, 41, 4)
    _compare("Connect4", 48, 100, 4, SERVER_DEFAULTS, boards=boards, turns=turns, moves=6)


@pytest.mark.parametrize("K", [1, 4])
def test_othello_score_utility(K):
    cfg = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False,
               score_utility_factor=0.15, score_scale=8.0)
    boards, turns = random_positions("Othello", 16, 30, 5)
    _compare("Othello", 16, 60, K, cfg, boards=boards, turns=turns, moves=3)


def test_othello_endgame_passes():
    cfg = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, use_symmetry=False,
               score_utility_factor=0.15, score_scale=6.0, value_decay=0.99)
    boards, turns = random_positions("Othello", 24, 58, 6)
    _compare("Othello", 24, 80, 4, cfg, boards=boards, turns=turns, moves=8)


def test_remove_all_vl_restores_tree():
    ref, orc = _pair("Connect4", 8)
    for e in (ref, orc):
        set_config(e, **SERVER_DEFAULTS)
    ev = ev_mod.HashEvaluator("Connect4", "hash")
    boards, turns = random_positions("Connect4", 8, 0, 0)
    for e in (ref, orc):
        playout(e, ev, boards, turns, 17, 4)
        e.search_batch_vl(4, boards, turns)
        e.remove_all_vl(4)
        e.remove_all_vl(4)          # idempotent
        playout(e, ev, boards, turns, 9, 4)
    assert np.array_equal(counts(ref, 8, 7), counts(orc, 8, 7))
    assert ref.get_all_root_stats().tobytes() == orc.get_all_root_stats().tobytes()


def test_env_random_games_match_reference_env():
    _, env_cpp = oracle.load_ref("parity")
    rng = np.random.default_rng(7)
    for game, sub in (("Connect4", env_cpp.connect4), ("Othello", env_cpp.othello)):
        for g in range(60):
            r, o = sub.Env(), oracle.OracleEnv(game)
            while True:
                assert np.array_equal(np.asarray(r.board).astype(np.int8), o.board)
                assert r.turn == o.turn and r.done() == o.done() and r.winPlayer() == o.winner()
                assert r.valid_move() == o.valid_moves()
                if r.done():
                    break
                mv = r.valid_move()
                a = mv[int(rng.integers(0, len(mv)))]
                r.step(a)
                o.step(a)
            # symmetry of the final position
            for s in range(sub.Env.NUM_SYMMETRIES):
                rs = r.apply_symmetry(s)
                os_ = oracle.OracleEnv(game)
                os_.import_board(o.board, o.turn)
                os_.apply_symmetry(s)
                assert np.array_equal(np.asarray(rs.board).astype(np.int8), os_.board)


def test_gomoku_restatement_vs_reference_env():
    """Byte-board Gomoku restatement vs the compiled reference Env (src/cpp/Gomoku.h): recorded rollouts replayed move by
    move, imports of unplayable boards (turn inference, full-scan winner), the D4 symmetries, and the three step errors."""
    _, ref_env = oracle.load_ref("parity")
    for size, k, games in ((15, 5, 12), (9, 4, 12), (6, 6, 6), (19, 5, 3), (32, 5, 1), (3, 3, 12)):
        for g in range(games):
            o = oracle.gomoku_rollout(size, k, 7, g)
            r = ref_env.gomoku.Env(size, k)
            e = oracle.OracleGomoku(size, k)
            for ply in range(o["plies"]):
                assert np.array_equal(np.asarray(r.board).astype(np.int8), o["boards"][ply]) and r.turn == o["turns"][ply]
                a = int(o["actions"][ply])
                mv = r.valid_move()
                assert mv == e.valid_moves() and a == mv[(oracle.lib().orc_rollout_hash(7, g, ply) >> 32) * len(mv) >> 32]
                r.step(a)
                assert e.step(a) == 0
                assert r.winPlayer() == o["winners"][ply] == e.winner() and r.done() == bool(o["dones"][ply]) == e.done()
            assert r.done() and np.array_equal(np.asarray(r.board).astype(np.int8), o["final"])
            assert e.step(0) == 1
            for s in range(8):
                e2 = oracle.OracleGomoku(size, k)
                e2.import_board(o["final"])
                e2.apply_symmetry(s)
                assert np.array_equal(np.asarray(r.apply_symmetry(s).board).astype(np.int8), e2.board)
    rng = np.random.default_rng(11)
    for size, k in ((8, 4), (15, 5), (5, 3)):
        for g in range(60):
            b = rng.choice(np.array([-1, 0, 1], np.int8), size=(size, size), p=(0.3 + 0.1 * (g % 3), 0.3, 0.4 - 0.1 * (g % 3)))
            r = ref_env.gomoku.Env(b.astype(np.float32), k)
            e = oracle.OracleGomoku(size, k)
            assert e.import_board(b) == 0
            assert (r.turn, r.done(), r.winPlayer()) == (e.turn, e.done(), e.winner())
    e = oracle.OracleGomoku(5, 3)
    assert e.step(-1) == 2 and e.step(25) == 2 and e.step(3) == 0 and e.step(3) == 3
    bad = np.zeros((5, 5), np.int8)
    bad[1, 1] = 3
    assert e.import_board(bad) == 7
