"""world_size-2 CPU (gloo) tests of the N>1 host logic: contiguous game sharding, sharding-invariant uids, and the
single all-gather of fixed-size packed trajectory records."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    lo, hi = sp_mod.shard_range(total, rank, world)
    rb = sp_mod.record_layout("Connect4").record_bytes
    cap = 8
    count = 3 + 2 * rank                                           # ragged: ranks finish different numbers of games
    local = torch.zeros((cap, rb), dtype=torch.uint8)
    for j in range(count):
        local[j, :] = (lo + j) % 251                               # recognisable payload
        local[j, 8:16] = torch.from_numpy(np.array([lo + j], np.uint64).view(np.uint8))
    gathered, counts = sp_mod.all_gather_records(local, count, cap)
    if rank == 0:
        out.put((lo, hi, counts, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 10
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    lo, hi, counts, g = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (lo, hi) == (0, 5) and counts == [3, 5] and g.shape[0] == 8
    uids = [int(r[8:16].view(np.uint64)[0]) for r in g]
    assert uids == [0, 1, 2, 5, 6, 7, 8, 9]                        # rank 0's games then rank 1's, in order


def test_shard_ranges_cover_everything():
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    for total in (7, 64, 65536):
        for world in (1, 2, 3, 8):
            r = [sp_mod.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
