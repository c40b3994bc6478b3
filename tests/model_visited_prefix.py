"""CPU model of the select / expansion traffic of the visited-prefix node layout against the current one, from the
C restatement's tree statistics on bench.py's workload (Connect4, n=200, K=4, server defaults, constant evaluator, fresh mid-game
roots).  Test infrastructure (it runs oracle/, so it lives under tests/, not tools/): it sizes next round's layout change
before any kernel is written.

  current layout : select gathers every edge slot of every node on the path      -> 32 B x edges scanned
                   an expansion writes one slot per legal move                   -> 32 B x edges created
  visited prefix : select gathers the 32-byte header + one record per child that -> 32 B x (nodes scanned + scanned edges whose
                   already has a record (allocated at its first visit)              child is allocated)
                   an expansion writes the header only; a first visit appends    -> 32 B x (expansions + first visits)
                   one record
python tests/model_visited_prefix.py [trees] [n_playout] [K]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    n_playout = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    A = 7
    boards, turns = bench.c4_random_roots(n, 1000)
    eng = oracle.OracleMCTS("Connect4", n)
    for k, v in bench.SERVER_DEFAULTS.items():
        setattr(eng.config, k, v)
    eng.config.c_base = 5.0 * n_playout                                  # server.py: c_base = 5 n (1000 at n = 200)
    eng.set_seed(0)
    bench.host_step(eng, boards, turns, n_playout, K, A, np.full(n, -1, np.int32))
    s = eng.tree_stats()
    sims = s["sims"]
    d, E, b = s["depth"] / sims, s["edges_scanned"] / sims, s["edges_created"] / sims
    alloc, seen, x = s["scanned_allocated"] / sims, s["scanned_visited"] / sims, s["expansions"] / sims
    first_visits = (s["nodes"] - n) / sims                       # every node except the roots was allocated by one first visit
    cur_sel, new_sel = 32 * E, 32 * (d + alloc)
    cur_exp, new_exp = 32 * b, 32 * (x + first_visits)
    print(f"trees {n}, simulations {sims}: per simulation  depth {d:.2f}  edges scanned {E:.2f}  of which allocated {alloc:.2f} / "
          f"visited {seen:.2f}  expansions {x:.2f}  edges created {b:.2f}  first visits {first_visits:.2f}")
    print(f"select gather   : current {cur_sel:7.1f} B   visited-prefix {new_sel:7.1f} B   ({new_sel / cur_sel:.2f}x)")
    print(f"expansion writes: current {cur_exp:7.1f} B   visited-prefix {new_exp:7.1f} B   ({new_exp / cur_exp:.2f}x)")
    print("per descent level (0 = root): nodes scanned / simulation, edges per node, allocated children per node, gather bytes now -> then")
    for l in range(8):
        ln, le, la = s["level_nodes"][l], s["level_edges"][l], s["level_allocated"][l]
        if ln:
            print(f"  level {l}{'+' if l == 7 else ' '}: {ln / sims:5.2f}  {le / ln:4.2f}  {la / ln:4.2f}   {32 * le / sims:6.1f} -> {32 * (ln + la) / sims:6.1f} B")
    xn, xr, xe = s["expanded_nodes"], s["expanded_revisited"], s["expanded_revisited_edges"]
    lazy = 32 * (x + xe / sims)                                   # header at expansion, the full block only on the second visit
    print(f"expanded nodes {xn / n:.1f} per tree, visited again {xr / n:.1f} ({100 * xr / xn:.0f} %): lazy full blocks (header at expansion, "
          f"block on the second visit) write {lazy:.1f} B per simulation ({lazy / cur_exp:.2f}x), select unchanged")
    print(f"select + expand : current {cur_sel + cur_exp:7.1f} B   visited-prefix {new_sel + new_exp:7.1f} B   "
          f"({(new_sel + new_exp) / (cur_sel + cur_exp):.2f}x)")


if __name__ == "__main__":
    main()
