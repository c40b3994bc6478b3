"""Multi-GPU: `torchrun` with 2 ranks, one per GPU - every rank self-plays its contiguous slot range and the
finished trajectories are all-gathered over NCCL inside the loop.  The union of what the ranks produced must equal, record for
record, what ONE GPU produces with all the slots (game i's result must not depend on the sharding), everything RNG-dependent on.
Needs >= 2 GPUs (`gpurun --gpus 2`); on a 1-GPU box the 1-GPU form of the same check runs in tests/test_gpu_selfplay.py."""
import importlib
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_self_play_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from harness import SERVER_DEFAULTS
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    G, npl, K, plies, every = 512, 40, 4, 30, 7
    out = str(tmp_path / "gathered.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "multirank_worker.py"), out, str(G), str(npl), str(K), str(plies), str(every)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = torch.load(out, weights_only=True)
    assert got["world"] == 2 and got["exchanges"] == (plies + every - 1) // every and got["bytes"] > 0
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3)
    sp = sp_mod.SelfPlay("Connect4", G, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=cfg, temperature=1.0, temp_decay_moves=8,
                         td_steps=4, seed=21, uid_base=0, uid_stride=G, out_capacity=6 * G)
    for _ in range(plies):
        sp.ply()
    whole = sp.drain().sorted_by_uid().cpu()
    assert len(whole) > G
    assert torch.equal(whole.games, got["games"]) and torch.equal(whole.pos, got["pos"])
