"""Host-side Env mirror (alphazero-al_b200/env_cpp) against the reference's pybind Env objects (oracle/_ref/parity,
when built), the C restatement, and the golden games recorded from the reference.  CPU only."""
import importlib
import os
import pickle

import numpy as np
import pytest

import oracle

env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")
HERE = os.path.dirname(os.path.abspath(__file__))
MODS = {"Connect4": env_cpp.connect4, "Othello": env_cpp.othello}


@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_env_reproduces_reference_golden_games(game):
    z = np.load(os.path.join(HERE, "golden", f"env_{game.lower()}_games.npz"))
    g = 0
    while f"g{g}_actions" in z:
        e = MODS[game].Env()
        for t, a in enumerate(z[f"g{g}_actions"]):
            assert e.board.dtype == np.float32 and np.array_equal(e.board.astype(np.int8), z[f"g{g}_boards"][t])
            assert np.array_equal(np.array(e.valid_mask(), dtype=np.uint8), z[f"g{g}_masks"][t])
            assert e.turn == z[f"g{g}_turns"][t]
            e.step(int(a))
            assert e.winPlayer() == z[f"g{g}_winners"][t] and e.done() == bool(z[f"g{g}_dones"][t])
        assert np.array_equal(e.board.astype(np.int8), z[f"g{g}_final"])
        g += 1
    assert g == 40


@pytest.mark.skipif(not oracle.ref_available("parity"), reason="oracle/_ref/parity not built")
@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_env_differential_vs_reference_objects(game):
    _, ref_env = oracle.load_ref("parity")
    rsub = ref_env.connect4 if game == "Connect4" else ref_env.othello
    msub = MODS[game]
    assert msub.Env.NUM_SYMMETRIES == rsub.Env.NUM_SYMMETRIES
    rng = np.random.default_rng(11)
    for g in range(40):
        r, m = rsub.Env(), msub.Env()
        ply = 0
        while True:
            assert np.array_equal(np.asarray(r.board), m.board) and r.turn == m.turn
            assert r.valid_move() == m.valid_move() and r.valid_mask() == m.valid_mask()
            assert r.done() == m.done() and r.winPlayer() == m.winPlayer() and r.check_full() == m.check_full()
            assert np.array_equal(r.current_state(), m.current_state()) and m.current_state().dtype == np.float32
            if ply % 5 == 0:
                for s in range(msub.Env.NUM_SYMMETRIES):
                    assert np.array_equal(np.asarray(r.apply_symmetry(s).board), m.apply_symmetry(s).board)
                    for a in m.valid_move()[:3]:
                        assert rsub.Env.inverse_symmetry_action(s, a) == msub.Env.inverse_symmetry_action(s, a)
                # from-board constructor + pickle round trip (turn is re-inferred from piece parity, env_common.h:69)
                r2, m2 = rsub.Env(np.asarray(r.board)), msub.Env(m.board)
                assert r2.turn == m2.turn and r2.valid_move() == m2.valid_move() and r2.winPlayer() == m2.winPlayer()
                r3, m3 = pickle.loads(pickle.dumps(r)), pickle.loads(pickle.dumps(m))
                assert r3.turn == m3.turn and np.array_equal(np.asarray(r3.board), m3.board) and r3.done() == m3.done()
            if r.done():
                break
            mv = r.valid_move()
            a = mv[int(rng.integers(0, len(mv)))]
            r.step(a)
            m.step(a)
            ply += 1
        c = m.copy()
        c.reset()
        assert m.done() and not c.done()            # copy is independent


@pytest.mark.skipif(not oracle.ref_available("parity"), reason="oracle/_ref/parity not built")
def test_gomoku_differential_vs_reference():
    _, ref_env = oracle.load_ref("parity")
    rng = np.random.default_rng(5)
    for size, k, games in ((15, 5, 8), (9, 5, 8), (6, 4, 8), (3, 3, 8), (19, 6, 2), (32, 5, 1), (2, 2, 2)):
        for g in range(games):
            r, m = ref_env.gomoku.Env(size, k), env_cpp.gomoku.Env(size, k)
            assert (r.board_size, r.rows, r.cols, r.n_in_row, r.action_size, r.num_symmetries) == \
                   (m.board_size, m.rows, m.cols, m.n_in_row, m.action_size, m.num_symmetries)
            ply = 0
            while not r.done():
                mv = r.valid_move()
                assert mv == m.valid_move()
                a = mv[int(rng.integers(0, len(mv)))]
                r.step(a)
                m.step(a)
                assert np.array_equal(np.asarray(r.board), m.board) and r.turn == m.turn
                assert r.done() == m.done() and r.winPlayer() == m.winPlayer() and r.check_full() == m.check_full()
                if ply % 7 == 0:
                    assert r.valid_mask() == m.valid_mask() and np.array_equal(r.current_state(), m.current_state())
                    for s in range(8):
                        assert np.array_equal(np.asarray(r.apply_symmetry(s).board), m.apply_symmetry(s).board)
                        assert r.inverse_symmetry_action(s, a) == m.inverse_symmetry_action(s, a)
                    r2, m2 = ref_env.gomoku.Env(np.asarray(r.board), k), env_cpp.gomoku.Env(m.board, k)
                    assert r2.turn == m2.turn and r2.done() == m2.done() and r2.winPlayer() == m2.winPlayer()
                    r3, m3 = pickle.loads(pickle.dumps(r)), pickle.loads(pickle.dumps(m))
                    assert r3.turn == m3.turn and np.array_equal(np.asarray(r3.board), m3.board) and r3.winPlayer() == m3.winPlayer()
                ply += 1
            for bad in (lambda e: e.step(a), lambda e: e.step(-1), lambda e: e.step(size * size)):
                for e in (r, m):
                    with pytest.raises(RuntimeError):
                        bad(e)
    assert ref_env.gomoku.Env(9, 5).coord_to_action(2, 3) == env_cpp.gomoku.Env(9, 5).coord_to_action(2, 3)
    assert tuple(ref_env.gomoku.Env(9, 5).action_to_coord(21)) == env_cpp.gomoku.Env(9, 5).action_to_coord(21)
    with pytest.raises(RuntimeError):
        env_cpp.gomoku.Env(4, 5)
    # imported boards nobody could have played: unequal stone counts, several lines of both colours (turn inference
    # Gomoku.h:194-199, first line in scan order wins :265-274), and cell values outside {-1, 0, 1}
    for size, k in ((8, 4), (15, 5), (5, 3)):
        for g in range(40):
            b = rng.choice(np.array([-1, 0, 1], np.float32), size=(size, size), p=(0.3 + 0.1 * (g % 3), 0.3, 0.4 - 0.1 * (g % 3)))
            r, m = ref_env.gomoku.Env(b, k), env_cpp.gomoku.Env(b, k)
            assert (r.turn, r.done(), r.winPlayer(), r.check_full()) == (m.turn, m.done(), m.winPlayer(), m.check_full())
            assert np.array_equal(np.asarray(r.board), m.board) and r.valid_move() == m.valid_move()
            for s in (1, 6):
                rs, ms = r.apply_symmetry(s), m.apply_symmetry(s)
                assert np.array_equal(np.asarray(rs.board), ms.board) and rs.winPlayer() == ms.winPlayer()
    bad = np.zeros((6, 6), np.float32)
    bad[2, 3] = 2
    for mod in (ref_env.gomoku, env_cpp.gomoku):
        with pytest.raises(RuntimeError):
            mod.Env(bad, 4)
        with pytest.raises(RuntimeError):
            mod.Env(6, 4).apply_symmetry(8)
        with pytest.raises(RuntimeError):
            mod.Env(6, 4).inverse_symmetry_action(-1, 0)
    with pytest.raises(RuntimeError):
        env_cpp.gomoku.Env(33, 5)                     # this implementation's limit (32-bit row masks)


def test_gomoku_basics_without_reference():
    e = env_cpp.gomoku.Env(7, 4)
    for a in (0, 7, 1, 8, 2, 9, 3):
        e.step(a)
    assert e.done() and e.winPlayer() == 1 and e.turn == -1
    with pytest.raises(RuntimeError):
        e.step(20)
    e2 = pickle.loads(pickle.dumps(e))
    assert e2.done() and e2.winPlayer() == 1 and np.array_equal(e2.board, e.board)


@pytest.mark.parametrize("size,k", [(15, 5), (8, 4), (32, 5), (4, 4)])
def test_gomoku_env_matches_restatement_on_rollouts(size, k):
    """env_cpp.gomoku.Env (row bit masks, C ABI host functions - the code the device kernels share) vs the byte-board
    restatement; runs without the compiled reference."""
    for g in range(6):
        o = oracle.gomoku_rollout(size, k, 9, g)
        m = env_cpp.gomoku.Env(size, k)
        for ply in range(o["plies"]):
            assert np.array_equal(m.board.astype(np.int8), o["boards"][ply]) and m.turn == o["turns"][ply]
            m.step(int(o["actions"][ply]))
            assert m.winPlayer() == o["winners"][ply] and m.done() == bool(o["dones"][ply])
        assert m.done() and np.array_equal(m.board.astype(np.int8), o["final"])
        for s in range(8):
            e = oracle.OracleGomoku(size, k)
            e.import_board(o["final"])
            e.apply_symmetry(s)
            ms = m.apply_symmetry(s)
            assert np.array_equal(ms.board.astype(np.int8), e.board)
            m2 = env_cpp.gomoku.Env(ms.board, k)               # fresh import of the transformed final board
            assert e.import_board(e.board) == 0
            assert (m2.turn, m2.done(), m2.winPlayer()) == (e.turn, e.done(), e.winner())


@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_env_matches_restatement_on_random_games(game):
    rng = np.random.default_rng(3)
    for g in range(30):
        m, o = MODS[game].Env(), oracle.OracleEnv(game)
        while not m.done():
            assert m.valid_move() == o.valid_moves() and np.array_equal(m.board.astype(np.int8), o.board)
            mv = m.valid_move()
            a = mv[int(rng.integers(0, len(mv)))]
            m.step(a)
            o.step(a)
        assert o.done() and m.winPlayer() == o.winner()
