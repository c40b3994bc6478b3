"""world_size-2 CPU (gloo) tests of the N>1 host logic: contiguous game sharding, sharding-invariant uids, and the single
all-gather of compact trajectory records: counts first, then max(count) rows per rank, ragged ranks, rank order."""
import importlib
import os
import socket

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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fixture_records import fixture_records, load_fixture
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    full = fixture_records(load_fixture("py_c4_selfplay_k4_sym"), "Connect4")
    m = len(full)
    lo, hi = (0, 7) if rank == 0 else (7, m)                        # ragged: ranks finished different numbers of games
    mine = sp_mod.Records.cat([sp_mod.Records("Connect4", full.games[i:i + 1], full.pos) for i in range(lo, hi)])
    got = sp_mod.all_gather_records(mine)
    empty = sp_mod.Records("Connect4", full.games[:0], full.pos[:0])
    got2 = sp_mod.all_gather_records(mine if rank == 1 else empty)  # a rank with nothing to contribute
    if rank == 0:
        out.put((torch.equal(got.games, full.games), torch.equal(got.pos, full.pos), len(got2), got2.uid.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_games, same_pos, n2, uids2 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_games and same_pos, "rank 0's games then rank 1's, positions re-based: the unsharded record set"
    assert n2 == 24 - 7 and uids2 == list(range(7, 24))


def test_shard_ranges_cover_everything():
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    for total in (7, 64, 65536):
        for world in (1, 2, 3, 8):
            r = [sp_mod.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
