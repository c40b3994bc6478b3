"""Device lockstep env kernels (config 2) against the C restatement: per-ply boards / masks / turns /
actions / winners / done flags bit-exact on a recorded slice, and a checksum of checksums over 1M games."""
import importlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
env_cpp = importlib.import_module("alphazero-al_b200.env_cpp")


@pytest.mark.parametrize("game,n_rec", [("Connect4", 4096), ("Othello", 1024)])
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
