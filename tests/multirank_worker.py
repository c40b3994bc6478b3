"""Worker of tests/test_gpu_multirank.py (one rank per GPU under torchrun): self-play on this rank's slot range, the NCCL
trajectory exchange inside the loop (on its side stream), rank 0 saves the gathered records sorted by uid."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path, G, npl, K, plies, every):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from harness import SERVER_DEFAULTS
    sp_mod = importlib.import_module("alphazero-al_b200.selfplay")
    ds = importlib.import_module("alphazero-al_b200.device_search")
    lo, hi = sp_mod.shard_range(G, rank, world)
    cfg = dict(SERVER_DEFAULTS, use_symmetry=True, dirichlet_alpha=0.3)
    sp = sp_mod.SelfPlay("Connect4", hi - lo, npl, K, ds.SyntheticEvaluator("Connect4", "hash"), search_cfg=cfg, temperature=1.0,
                         temp_decay_moves=8, td_steps=4, seed=21, uid_base=lo, uid_stride=G, device=local, out_capacity=6 * (hi - lo))
    exch = sp_mod.TrajectoryExchange("Connect4", sp.out_capacity, dev)
    parts, pending = [], None
    for p in range(plies):
        sp.ply()
        if (p + 1) % every == 0 or p + 1 == plies:
            ring = sp.hand_over()
            if pending is not None:
                exch.submit(pending)
            pending = ring
    exch.submit(pending)
    while exch.pending:
        rec = exch.collect()
        if len(rec):
            parts.append(rec)
    allrec = sp_mod.Records.cat(parts).sorted_by_uid()
    if rank == 0:
        torch.save({"games": allrec.games.cpu(), "pos": allrec.pos.cpu(), "world": world, "exchanges": len(exch.ms),
                    "bytes": exch.bytes_gathered}, out_path)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], *(int(x) for x in sys.argv[2:7]))
