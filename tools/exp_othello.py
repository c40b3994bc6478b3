"""Othello config-4 search (N trees, n=400, K=4, score utility) for ncu / timing.  python tools/exp_othello.py [n] [lanes]"""
import importlib, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
oth = dict(c_init=1.4, c_base=2000.0, fpu_reduction=0.2, dirichlet_alpha=0.3, noise_epsilon=0.25, use_symmetry=True,
           score_utility_factor=0.15, score_scale=8.0)
bc.run("Othello", n, 400, 4, oth, "hash", steps=2, label=f"Othello N={n}", lanes=lanes)
