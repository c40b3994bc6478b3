"""Shared drivers for the parity tests: they replay the reference wrapper's playout loop
(src/MCTS_cpp.py:89-359: one non-VL warm-up simulation, then ceil((n-1)/K) virtual-loss iterations with
cur_K = min(K, remaining)) on any engine exposing the mcts_cpp.BatchedMCTS_<Game> methods."""
import numpy as np


def playout(engine, evaluator, boards, turns, n_playout, K, record=None):
    """Run n_playout simulations per tree.  `record` (a list) receives every leaf tuple for comparison.  An evaluator with a true
    `wants_mask` attribute also receives the leaves' legal masks (what the reference wrapper hands to predict())."""
    boards = np.ascontiguousarray(boards, dtype=np.int8)
    turns = np.ascontiguousarray(turns, dtype=np.int32)
    if getattr(evaluator, "wants_mask", False):
        inner = evaluator

        class _WithMask:
            def __call__(self, lb, lt, it, td, tp1, tp2):
                return inner(lb, lt, it, td, tp1, tp2, self.vm)
        evaluator = _WithMask()
        base_engine = engine

        class _Tap:
            def __getattr__(self, name):
                f = getattr(base_engine, name)
                if name in ("search_batch", "search_batch_vl"):
                    def g(*a):
                        out = f(*a)
                        evaluator.vm = out[-1]
                        return out
                    return g
                return f
        engine = _Tap()
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


def compare_engines(ea, eb, game, n, n_playout, K, cfg, mode="hash", boards=None, turns=None, moves=1,
                    compare_leaves=True, seed=None):
    """Drive two engines (same mcts_cpp surface) through identical playouts + tree reuse and require identical
    leaves (optional), visit counts and root statistics, bit for bit.  Returns the last visit counts."""
    import importlib
    import oracle
    ev_mod = importlib.import_module("alphazero-al_b200.evaluators")
    A = oracle.ACTION_SIZE[game]
    set_config(ea, **cfg)
    set_config(eb, **cfg)
    if seed is not None:
        ea.set_seed(seed)
        eb.set_seed(seed)
    ev = ev_mod.HashEvaluator(game, mode)
    if boards is None:
        boards, turns = random_positions(game, n, 0, 0)
    envs = [oracle.OracleEnv(game) for _ in range(n)]
    for i, e in enumerate(envs):
        e.import_board(boards[i], turns[i])
    c1 = None
    for mv in range(moves):
        b = np.stack([e.board for e in envs])
        t = np.array([e.turn for e in envs], np.int32)
        r1, r2 = [], []
        playout(ea, ev, b, t, n_playout, K, r1)
        playout(eb, ev, b, t, n_playout, K, r2)
        if compare_leaves:
            assert len(r1) == len(r2)
            for it, (x, y) in enumerate(zip(r1, r2)):
                for j, (u, v) in enumerate(zip(x, y)):
                    if not np.array_equal(u, v):
                        bad = np.where((u != v).reshape(len(u), -1).any(axis=1))[0]
                        raise AssertionError(f"move {mv} iteration {it} output {j} differs in rows {bad[:8]} "
                                             f"(of {len(bad)}): a={u[bad[0]]!r} b={v[bad[0]]!r}")
        c1, c2 = counts(ea, n, A), counts(eb, n, A)
        if not np.array_equal(c1, c2):
            bad = np.where((c1 != c2).any(axis=1))[0]
            raise AssertionError(f"move {mv}: visit counts differ in {len(bad)} trees, first {bad[0]}: {c1[bad[0]]} vs {c2[bad[0]]}")
        s1, s2 = ea.get_all_root_stats(), eb.get_all_root_stats()
        if s1.tobytes() != s2.tobytes():
            bad = np.where((s1.view(np.uint32) != s2.view(np.uint32)).any(axis=1))[0]
            cols = np.where(s1.view(np.uint32)[bad[0]] != s2.view(np.uint32)[bad[0]])[0]
            raise AssertionError(f"move {mv}: root stats differ in {len(bad)} trees; tree {bad[0]} cols {cols[:10]}: "
                                 f"{s1[bad[0]][cols[:10]]} vs {s2[bad[0]][cols[:10]]}")
        # play the most visited action (ties -> lowest); finished / unsearchable games restart from scratch
        acts = np.zeros(n, np.int32)
        for i, e in enumerate(envs):
            if e.done() or c1[i].sum() == 0:
                e.reset()
                ea.reset_env(i)
                eb.reset_env(i)
                acts[i] = -1
            else:
                acts[i] = int(np.argmax(c1[i]))
                e.step(acts[i])
        ea.prune_roots(acts)
        eb.prune_roots(acts)
    return c1
