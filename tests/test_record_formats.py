"""Compact trajectory records <-> the reference's formats, on fixtures the UNMODIFIED reference produced
(tests/golden/py_*_selfplay_*: Game.batch_self_play tuples, the ReplayBuffer.save file, the actor's upload pickle).

CPU: record container logic (header views, sort, concatenation).  GPU: the expansion kernel reproduces every training tuple of the
fixture, `save_replay_pt` writes the file the reference's ReplayBuffer wrote, `to_upload_payload` the pickle its actor uploads; in
the build container the reference's own `ReplayBuffer.load` additionally reads our file."""
import importlib
import json
import os
import pickle

import numpy as np
import pytest
import torch

from fixture_records import GOLD, fixture_records, load_fixture

sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
CASES = json.load(open(os.path.join(GOLD, "py_cases.json")))["selfplay"]
NAMES = list(CASES)


def test_record_container_sort_and_cat():
    z = load_fixture("py_c4_selfplay_k4_sym")
    rec = fixture_records(z, "Connect4", uid0=100)
    m = len(rec)
    assert m == len(z["length"]) and rec.positions == int(z["length"].sum())
    assert rec.uid.tolist() == list(range(100, 100 + m)) and rec.length.tolist() == z["length"].tolist() and rec.winner.tolist() == z["winner"].tolist()
    perm = torch.randperm(m, generator=torch.Generator().manual_seed(1))
    parts = [sp_mod.Records("Connect4", rec.games[i:i + 1], rec.pos) for i in perm.tolist()]     # each part drags the whole position array along
    shuffled = sp_mod.Records.cat(parts)
    assert shuffled.positions == rec.positions and shuffled.uid.tolist() == (perm + 100).tolist()
    back = shuffled.sorted_by_uid()
    assert torch.equal(back.games, rec.games) and torch.equal(back.pos, rec.pos)


def _assert_tensors_match_fixture(t, z, td):
    n = lambda k: t[k].cpu().numpy()
    assert np.array_equal(n("state"), z["state"]) and n("state").dtype == np.int8
    assert n("prob").tobytes() == z["prob"].tobytes() and n("root_wdl").tobytes() == z["root_wdl"].tobytes()
    assert np.array_equal(n("winner")[:, 0], z["winner_z"]) and np.array_equal(n("steps_to_end")[:, 0], z["steps_to_end"])
    assert np.array_equal(n("aux_target")[:, 0], z["aux"]) and np.array_equal(n("valid_mask"), z["valid_mask"])
    if td > 0:
        assert n("future_root_wdl").tobytes() == z["future_root_wdl"].tobytes()
    else:
        assert not n("future_root_wdl").any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_expand_reproduces_reference_tuples_pt_and_pickle(name, tmp_path):
    from test_gpu_dropin_python_layer import assert_pt_equal, assert_same_structure
    case = CASES[name]
    game, td = case["game"], case["td_steps"]
    z = load_fixture(name)
    rec = fixture_records(z, game).to("cuda")
    t = rec.to_replay_tensors(td)
    _assert_tensors_match_fixture(t, z, td)
    games = rec.unpack(td)
    assert [g["winner"] for g in games] == z["winner"].tolist() and [g["length"] for g in games] == z["length"].tolist()
    assert len(games[0]["tuples"][1][0]) == int(z["tuple_width"])
    if case.get("formats"):
        p = str(tmp_path / "buffer.pt")
        sp_mod.save_replay_pt(p, t)
        assert_pt_equal(p, os.path.join(GOLD, name + ".pt"))
        with open(os.path.join(GOLD, name + ".pkl"), "rb") as f:
            assert_same_structure(pickle.loads(sp_mod.to_upload_payload(games)), pickle.load(f))


@pytest.mark.gpu
def test_reference_replay_buffer_loads_our_file(tmp_path):
    """Build container / GPU box with oracle/_ref/pysrc: the reference's own ReplayBuffer.load (src/ReplayBuffer.py:40-62) reads
    the file save_replay_pt wrote and ends up with the fixture's contents."""
    from oracle import refstack
    if not refstack.available("parity"):
        pytest.skip("oracle/_ref/pysrc not present")
    name = "py_c4_selfplay_k4_sym"
    z = load_fixture(name)
    t = fixture_records(z, "Connect4").to("cuda").to_replay_tensors(CASES[name]["td_steps"])
    p = str(tmp_path / "ours.pt")
    sp_mod.save_replay_pt(p, t)
    ov = refstack.make_overlay(str(tmp_path / "ref"), "reference")
    out = str(tmp_path / "loaded.npz")
    refstack.run_driver(ov, "load_pt", out, dict(path=p, rows=int(z["length"].sum()), A=7, R=6, C=7))
    got = np.load(out)
    assert int(got["ptr"]) == int(z["length"].sum()) and int(got["len"]) == int(z["length"].sum())
    assert np.array_equal(got["state"], z["state"]) and got["prob"].tobytes() == z["prob"].tobytes()
    assert np.array_equal(got["winner"][:, 0], z["winner_z"]) and np.array_equal(got["valid_mask"], z["valid_mask"])
    assert got["future_root_wdl"].tobytes() == z["future_root_wdl"].tobytes()
