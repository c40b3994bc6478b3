"""Shared drivers for the parity tests: they replay the reference wrapper's playout loop
(src/MCTS_cpp.py:89-359: one non-VL warm-up simulation, then ceil((n-1)/K) virtual-loss iterations with
cur_K = min(K, remaining)) on any engine exposing the mcts_cpp.BatchedMCTS_<Game> methods."""
import numpy as np


def playout(engine, evaluator, boards, turns, n_playout, K, record=None):
    """Run n_playout simulations per tree.  `record` (a list) receives every leaf tuple for comparison."""
    boards = np.ascontiguousarray(boards, dtype=np.int8)
    turns = np.ascontiguousarray(turns, dtype=np.int32)
    if K <= 1:
        for _ in range(n_playout):
            lb, td, tp1, tp2, it, lt, vm = engine.search_batch(boards, turns)
            if record is not None:
                record.append((lb.copy(), td.copy(), tp1.copy(), tp2.copy(), it.copy(), lt.copy(), vm.copy()))
            probs, d, p1w, p2w, ml = evaluator(lb, lt, it, td, tp1, tp2)
            engine.backprop_batch(probs, d, p1w, p2w, ml, it)
        return
    remaining = n_playout
    if remaining > 0:
        lb, td, tp1, tp2, it, lt, vm = engine.search_batch(boards, turns)
        if record is not None:
            record.append((lb.copy(), td.copy(), tp1.copy(), tp2.copy(), it.copy(), lt.copy(), vm.copy()))
        probs, d, p1w, p2w, ml = evaluator(lb, lt, it, td, tp1, tp2)
        engine.backprop_batch(probs, d, p1w, p2w, ml, it)
        remaining -= 1
    while remaining > 0:
        cur = min(K, remaining)
        remaining -= cur
        lb, td, tp1, tp2, it, lt, sym, vm = engine.search_batch_vl(cur, boards, turns)
        if record is not None:
            record.append((lb.copy(), td.copy(), tp1.copy(), tp2.copy(), it.copy(), lt.copy(), sym.copy(), vm.copy()))
        probs, d, p1w, p2w, ml = evaluator(lb, lt, it, td, tp1, tp2)
        engine.backprop_batch_vl(cur, probs, d, p1w, p2w, ml, it, sym)


def counts(engine, n, A):
    return np.asarray(engine.get_all_counts(), dtype=np.int64).reshape(n, A)


def set_config(engine, **kw):
    cfg = engine.config
    for k, v in kw.items():
        setattr(cfg, k, v)


SERVER_DEFAULTS = dict(c_init=1.4, c_base=1000.0, fpu_reduction=0.2, dirichlet_alpha=0.0, noise_epsilon=0.25,
                       mlh_slope=0.1, mlh_cap=0.2, use_symmetry=False, value_decay=1.0)


def random_positions(game, n, max_plies, seed):
    """Mid-game roots from uniformly random legal playouts on the C restatement (never terminal)."""
    from oracle import OracleEnv, BOARD_SHAPE
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, *BOARD_SHAPE[game]), np.int8)
    turns = np.ones(n, np.int32)
    for i in range(n):
        while True:
            e = OracleEnv(game)
            plies = int(rng.integers(0, max_plies + 1))
            ok = True
            for _ in range(plies):
                mv = e.valid_moves()
                e.step(mv[int(rng.integers(0, len(mv)))])
                if e.done():
                    ok = False
                    break
            if ok:
                boards[i], turns[i] = e.board, e.turn
                break
    return boards, turns
