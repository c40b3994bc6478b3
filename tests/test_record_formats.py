"""Packed trajectory records -> the reference's replay-buffer tensors, .pt file and upload pickle (CPU only)."""
import importlib
import os
import pickle
import tempfile

import numpy as np
import torch

sp_mod = importlib.import_module("alphazero-al_b200.selfplay")


def _fake_records(game, lengths, seed=0):
    """Builds packed records on the host with the layout az_selfplay_layout_for reports."""
    rng = np.random.default_rng(seed)
    L = sp_mod.record_layout(game)
    gid, R, Cc, A, T = sp_mod._G[game]
    S = R * Cc
    out = np.zeros((len(lengths), L.record_bytes), np.uint8)
    truth = []
    for i, n in enumerate(lengths):
        rec = out[i]
        rec[L.off_header:L.off_header + 8].view(np.int32)[:] = (n, 1 - 2 * (i % 2))
        rec[L.off_header + 8:L.off_header + 16].view(np.uint64)[:] = 1000 + i
        st = rng.integers(-1, 2, size=(n, 3, R, Cc)).astype(np.int8)
        pr = rng.random((n, A)).astype(np.float32)
        rw, fw = rng.random((n, 3)).astype(np.float32), rng.random((n, 3)).astype(np.float32)
        wz = np.full(n, 1 - 2 * (i % 2), np.int8)
        ste = np.arange(n - 1, -1, -1).astype(np.int16)
        aux = rng.integers(-60, 60, size=n).astype(np.int16)
        mk = rng.integers(0, 2, size=(n, A)).astype(np.uint8)
        rec[L.off_state:L.off_state + n * 3 * S] = st.reshape(-1).view(np.uint8)
        rec[L.off_prob:L.off_prob + n * A * 4] = pr.reshape(-1).view(np.uint8)
        rec[L.off_root_wdl:L.off_root_wdl + n * 12] = rw.reshape(-1).view(np.uint8)
        rec[L.off_future:L.off_future + n * 12] = fw.reshape(-1).view(np.uint8)
        rec[L.off_winner:L.off_winner + n] = wz.view(np.uint8)
        rec[L.off_steps:L.off_steps + n * 2] = ste.view(np.uint8)
        rec[L.off_aux:L.off_aux + n * 2] = aux.view(np.uint8)
        rec[L.off_mask:L.off_mask + n * A] = mk.reshape(-1)
        truth.append(dict(state=st, prob=pr, root_wdl=rw, future_root_wdl=fw, winner=wz, steps_to_end=ste, aux=aux, mask=mk.astype(bool)))
    return out, truth


def test_replay_tensors_pt_and_upload_payload():
    for game, lengths in (("Connect4", [8, 43, 1, 22]), ("Othello", [61, 5])):
        packed, truth = _fake_records(game, lengths)
        t = sp_mod.to_replay_tensors(torch.from_numpy(packed), game)
        n = sum(lengths)
        gid, R, Cc, A, T = sp_mod._G[game]
        assert t["state"].shape == (n, 3, R, Cc) and t["state"].dtype == torch.int8
        assert t["prob"].shape == (n, A) and t["winner"].shape == (n, 1) and t["winner"].dtype == torch.int8
        assert t["steps_to_end"].dtype == torch.int16 and t["aux_target"].dtype == torch.int16
        assert t["valid_mask"].dtype == torch.bool and t["future_root_wdl"].shape == (n, 3)
        cat = lambda k: np.concatenate([x[k] for x in truth])
        assert np.array_equal(t["state"].numpy(), cat("state")) and np.array_equal(t["prob"].numpy(), cat("prob"))
        assert np.array_equal(t["root_wdl"].numpy(), cat("root_wdl")) and np.array_equal(t["future_root_wdl"].numpy(), cat("future_root_wdl"))
        assert np.array_equal(t["winner"].numpy()[:, 0], cat("winner")) and np.array_equal(t["steps_to_end"].numpy()[:, 0], cat("steps_to_end"))
        assert np.array_equal(t["aux_target"].numpy()[:, 0], cat("aux")) and np.array_equal(t["valid_mask"].numpy(), cat("mask"))
        games = sp_mod.unpack_records(packed, game, td_steps=3)
        assert [g["uid"] for g in games] == [1000 + i for i in range(len(lengths))]
        assert [len(g["tuples"][1]) for g in games] == lengths and len(games[0]["tuples"][1][0]) == 8
        payload = pickle.loads(sp_mod.to_upload_payload(games))
        assert payload["__az__"] is True and len(payload["data"]) == len(lengths) and len(payload["data"][1]) == lengths[1]
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "buffer.pt")
            sp_mod.save_replay_pt(p, t)
            sd = torch.load(p, weights_only=True)
            assert sd["_ptr"] == n and sd["current_capacity"] == n and set(sd) >= {"state", "prob", "winner", "steps_to_end", "aux_target",
                                                                                 "root_wdl", "valid_mask", "future_root_wdl"}
