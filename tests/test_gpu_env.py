"""Device lockstep env kernels (config 2) against the C restatement: per-ply boards / masks / turns /
actions / winners / done flags bit-exact on a recorded slice, and a checksum of checksums over 1M games."""
import importlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")


@pytest.mark.parametrize("game,n_rec", [("Connect4", 10240), ("Othello", 1024)])
def test_recorded_rollouts_match_restatement_per_ply(game, n_rec):
    import torch
    be = env_cpp.BatchedEnv(game, n_rec)
    digest, plies, rec = be.random_rollouts(seed=0, first_game=0, n_record=n_rec)
    torch.cuda.synchronize()
    rec = {k: v.cpu().numpy() for k, v in rec.items()}
    plies, digest = plies.cpu().numpy(), digest.cpu().numpy().view(np.uint64)
    for g in range(n_rec):
        o = oracle.env_rollout(game, 0, g)
        n = o["plies"]
        assert plies[g] == n and digest[g] == o["digest"]
        assert np.array_equal(rec["boards"][g, :n], o["boards"]) and np.array_equal(rec["masks"][g, :n], o["masks"])
        assert np.array_equal(rec["turns"][g, :n], o["turns"]) and np.array_equal(rec["actions"][g, :n], o["actions"])
        assert np.array_equal(rec["winners"][g, :n], o["winners"]) and np.array_equal(rec["dones"][g, :n], o["dones"])


@pytest.mark.parametrize("game,n_rec", [("Connect4", 10240), ("Othello", 512)])
def test_recorded_rollouts_match_the_compiled_reference_env_per_ply(game, n_rec):
    """The same recorded slice against the UNMODIFIED reference itself: every game is replayed, action by action, on an
    `env_cpp.<game>.Env` object of the compiled reference (oracle/_ref/parity travels to the GPU box) and its board, legal mask, side
    to move, winner and done flag are compared with what the device kernels recorded at every ply (BASELINE config 2's check)."""
    import torch
    if not oracle.ref_available("parity"):
        pytest.skip("oracle/_ref/parity not present")
    _, ref_env = oracle.load_ref("parity")
    sub = getattr(ref_env, game.lower())
    be = env_cpp.BatchedEnv(game, n_rec)
    digest, plies, rec = be.random_rollouts(seed=0, first_game=0, n_record=n_rec)
    torch.cuda.synchronize()
    rec = {k: v.cpu().numpy() for k, v in rec.items()}
    plies = plies.cpu().numpy()
    A = be.A
    for g in range(n_rec):
        e = sub.Env()
        n = int(plies[g])
        for t in range(n):
            assert np.array_equal(np.asarray(e.board).astype(np.int8), rec["boards"][g, t]), (g, t)
            assert np.array_equal(np.asarray(e.valid_mask(), dtype=np.uint8), rec["masks"][g, t][:A]) and e.turn == rec["turns"][g, t]
            e.step(int(rec["actions"][g, t]))
            assert e.winPlayer() == rec["winners"][g, t] and bool(e.done()) == bool(rec["dones"][g, t]), (g, t)
        assert e.done()


def test_one_million_connect4_games_checksum():
    import torch
    n = 1_000_000
    be = env_cpp.BatchedEnv("Connect4", n)
    digest, plies, _ = be.random_rollouts(seed=0, first_game=0)
    torch.cuda.synchronize()
    d = digest.cpu().numpy().view(np.uint64)
    p = plies.cpu().numpy()
    od, op = oracle.env_rollout_digests("Connect4", 0, 0, n)
    assert np.array_equal(p, op)
    assert np.array_equal(d, od)
    assert int(np.bitwise_xor.reduce(d)) == int(np.bitwise_xor.reduce(od))
    assert 7 <= p.min() and p.max() <= 42


@pytest.mark.parametrize("game", ["Connect4", "Othello"])
def test_lockstep_step_and_observe(game):
    import torch
    n = 512
    be = env_cpp.BatchedEnv(game, n)
    envs = [oracle.OracleEnv(game) for _ in range(n)]
    rng = np.random.default_rng(1)
    for ply in range(70):
        obs = {k: (v.cpu().numpy() if v is not None else None) for k, v in be.observe().items()}
        acts = np.full(n, -1, np.int32)
        for i, e in enumerate(envs):
            assert np.array_equal(obs["boards"][i], e.board) and obs["turns"][i] == e.turn
            assert bool(obs["dones"][i]) == e.done() and obs["winners"][i] == e.winner()
            mask = np.zeros(be.A, np.uint8)
            mask[e.valid_moves()] = 1
            assert np.array_equal(obs["masks"][i], mask)
            if not e.done():
                mv = e.valid_moves()
                acts[i] = mv[int(rng.integers(0, len(mv)))]
                e.step(int(acts[i]))
        if (acts < 0).all():
            break
        be.step(torch.from_numpy(acts).to(be.device))
    assert all(e.done() for e in envs)


def test_env_states_feed_the_search_directly():
    """az_root records produced by the env kernels are valid search roots (no byte boards in between)."""
    import torch
    from harness import SERVER_DEFAULTS, counts, playout, set_config
    mcts_cpp = importlib.import_module("alphazero-al_b200.mcts_cpp")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    n = 256
    be = env_cpp.BatchedEnv("Connect4", n)
    rng = np.random.default_rng(2)
    for _ in range(6):
        m = be.observe()["masks"].cpu().numpy()
        acts = np.array([rng.choice(np.nonzero(r)[0]) for r in m], np.int32)
        be.step(torch.from_numpy(acts).to(be.device))
    obs = be.observe()
    a, b = mcts_cpp.BatchedMCTS_Connect4(n), mcts_cpp.BatchedMCTS_Connect4(n)
    for e in (a, b):
        set_config(e, **SERVER_DEFAULTS)
    playout(a, ev_mod.HashEvaluator("Connect4", "hash"), obs["boards"].cpu().numpy(), obs["turns"].cpu().numpy(), 60, 4)
    buf = ds.LeafBuffers(n, n * 4, 7, (6, 7), be.device)
    buf.roots.copy_(be.states)
    ds.playout_device(b, buf, 60, 4, ds.SyntheticEvaluator("Connect4", "hash"))
    torch.cuda.synchronize()
    assert np.array_equal(counts(a, n, 7), counts(b, n, 7))


# ---- Gomoku (Env-only in the reference, src/cpp/Gomoku.h): lockstep device kernels vs the byte-board restatement ----
@pytest.mark.parametrize("size,k,n_rec", [(15, 5, 192), (9, 4, 256), (32, 5, 8), (3, 3, 64), (6, 6, 64)])
def test_gomoku_recorded_rollouts_match_restatement_per_ply(size, k, n_rec):
    import torch
    be = env_cpp.BatchedGomoku(n_rec + 37, size, k)              # ragged tail: 37 unrecorded games
    digest, plies, rec = be.random_rollouts(seed=3, first_game=5, n_record=n_rec)
    torch.cuda.synchronize()
    rec = {kk: v.cpu().numpy() for kk, v in rec.items()}
    plies, digest = plies.cpu().numpy(), digest.cpu().numpy().view(np.uint64)
    obs = {kk: v.cpu().numpy() for kk, v in be.observe().items()}
    for g in range(n_rec + 37):
        o = oracle.gomoku_rollout(size, k, 3, 5 + g, record=g < n_rec)
        n = o["plies"]
        assert plies[g] == n and digest[g] == o["digest"]
        assert np.array_equal(obs["boards"][g], o["final"]) and obs["dones"][g] == 1
        assert np.array_equal(obs["masks"][g], (o["final"].reshape(-1) == 0).astype(np.uint8))
        if g < n_rec:
            assert np.array_equal(rec["boards"][g, :n], o["boards"]) and np.array_equal(rec["turns"][g, :n], o["turns"])
            assert np.array_equal(rec["actions"][g, :n], o["actions"]) and np.array_equal(rec["winners"][g, :n], o["winners"])
            assert np.array_equal(rec["dones"][g, :n], o["dones"]) and obs["winners"][g] == o["winners"][-1]


def test_gomoku_128k_games_checksum():
    import torch
    n = 131_072
    be = env_cpp.BatchedGomoku(n, 15, 5)
    digest, plies, _ = be.random_rollouts(seed=1, first_game=0, keep_final=False)
    torch.cuda.synchronize()
    d, p = digest.cpu().numpy().view(np.uint64), plies.cpu().numpy()
    od, op = oracle.gomoku_rollout_digests(15, 5, 1, 0, n)
    assert np.array_equal(p, op) and np.array_equal(d, od)
    assert int(np.bitwise_xor.reduce(d)) == int(np.bitwise_xor.reduce(od))
    assert 9 <= p.min() and p.max() <= 225


@pytest.mark.parametrize("size,k", [(15, 5), (7, 4), (32, 6)])
def test_gomoku_lockstep_step_observe_symmetry(size, k):
    import torch
    n = 200
    be = env_cpp.BatchedGomoku(n, size, k)
    envs = [oracle.OracleGomoku(size, k) for _ in range(n)]
    rng = np.random.default_rng(4)
    status = torch.zeros(n, dtype=torch.uint8, device=be.device)
    winners = torch.zeros(n, dtype=torch.int32, device=be.device)
    dones = torch.zeros(n, dtype=torch.uint8, device=be.device)
    for ply in range(size * size + 2):
        obs = {kk: v.cpu().numpy() for kk, v in be.observe().items()}
        acts = np.full(n, -1, np.int32)
        want = np.zeros(n, np.uint8)
        for i, e in enumerate(envs):
            assert np.array_equal(obs["boards"][i], e.board) and obs["turns"][i] == e.turn
            assert bool(obs["dones"][i]) == e.done() and obs["winners"][i] == e.winner()
            if e.done():
                acts[i] = 0 if i % 2 else -1                     # finished games are skipped whatever the action
                continue
            mv = e.valid_moves()
            roll = rng.random()
            if roll < 0.05 and e.n_pieces:                        # occupied cell: reported, game untouched
                acts[i], want[i] = e.last_action, 3
            elif roll < 0.08:                                     # out of range
                acts[i], want[i] = size * size + int(rng.integers(0, 5)), 2
            else:
                acts[i] = mv[int(rng.integers(0, len(mv)))]
                assert e.step(int(acts[i])) == 0
        if all(e.done() for e in envs):
            break
        be.step(torch.from_numpy(acts).to(be.device), status, winners, dones)
        assert np.array_equal(status.cpu().numpy(), want)
        assert np.array_equal(winners.cpu().numpy(), [e.winner() for e in envs])
        assert np.array_equal(dones.cpu().numpy(), [int(e.done()) for e in envs])
        if ply % 16 == 5:                                         # D4 symmetries on the device (ids outside 1..7 = no-op)
            syms = rng.integers(-1, 9, n).astype(np.int32)
            be.apply_symmetry(torch.from_numpy(syms).to(be.device))
            for e, s in zip(envs, syms):
                e.apply_symmetry(int(s))
    assert all(e.done() for e in envs)
