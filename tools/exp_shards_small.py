"""Do independent tree shards on their own streams help the small-batch configurations too?  (BASELINE config 3: Connect4 8192
trees n=800 K=8; config 4: Othello 4096 trees n=400 K=4; auto_shards() keeps batches below 16 384 trees whole.)
    AZB200_SHARDS=k python tools/exp_shards_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc
c4 = dict(c_init=1.4, c_base=4000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, mlh_slope=0.1, mlh_cap=0.2, use_symmetry=True)
oth = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, use_symmetry=True,
           score_utility_factor=0.15, score_scale=8.0)
tag = "shards=" + os.environ.get("AZB200_SHARDS", "auto")
bc.run("Connect4", 8192, 800, 8, c4, "equivariant", steps=3, label="config 3 " + tag)
bc.run("Othello", 4096, 400, 4, oth, "hash", steps=3, label="config 4 " + tag)
bc.run("Othello", 16384, 400, 4, oth, "hash", steps=2, label="Othello 16384 " + tag)
