"""CPU-only checks of the device playout loop's host logic (alphazero-al_b200/device_search.py): it must schedule exactly
the iterations of the reference wrapper (src/MCTS_cpp.py:217-357: one non-VL warm-up simulation, then ceil((n-1)/K)
virtual-loss iterations with cur_K = min(K, remaining)), and pick sensible shard counts."""
import importlib

import pytest
import torch

ds = importlib.import_module("alphazero-al_b200.device_search")


class _FakeEngine:
    def __init__(self, n):
        self.n, self.calls = n, []

    def get_num_envs(self):
        return self.n

    def search_dev(self, k, roots, leaves, stream):
        self.calls.append(("search", k))

    def backprop_dev(self, k, *a):
        self.calls.append(("backprop", k))


class _Buf:
    def __init__(self, n, rows):
        self.rows = rows
        self.roots = torch.zeros((n, 32), dtype=torch.uint8)
        self.leaves = torch.zeros((rows, 32), dtype=torch.uint8)
        self.policy = torch.zeros((rows, 7))
        self.d = self.p1w = self.p2w = self.ml = torch.zeros(rows)


@pytest.mark.parametrize("n_playout,K,expect", [
    (200, 4, [0] + [4] * 49 + [3]),          # the BASELINE configuration: 1 + 49*4 + 3 = 200
    (800, 8, [0] + [8] * 99 + [7]),
    (9, 4, [0, 4, 4]),
    (1, 4, [0]),
    (0, 4, []),
    (5, 1, [0] * 5),                         # vl_batch <= 1: n_playout non-VL simulations
])
def test_iteration_schedule_is_the_reference_wrappers(n_playout, K, expect):
    eng, rows_seen = _FakeEngine(96), []
    buf = _Buf(96, 96 * max(K, 1))
    launches = ds.playout_device(eng, buf, n_playout, K, lambda b, rows, s, row0=0: rows_seen.append(rows), stream=1, shards=1)
    assert [k for what, k in eng.calls if what == "search"] == expect
    assert [k for what, k in eng.calls if what == "backprop"] == expect
    assert rows_seen == [96 * max(k, 1) for k in expect]
    assert launches == 3 * len(expect)
    assert sum(max(k, 1) for k in expect) == n_playout


def test_auto_shards(monkeypatch):
    monkeypatch.delenv("AZB200_SHARDS", raising=False)
    assert ds.auto_shards(100) == 1 and ds.auto_shards(8192) == 1
    assert ds.auto_shards(16384) == 2 and ds.auto_shards(32768) == 4 and ds.auto_shards(65536) == 8 and ds.auto_shards(1 << 20) == 8
    monkeypatch.setenv("AZB200_SHARDS", "3")
    assert ds.auto_shards(65536) == 3


def test_network_batch_buckets():
    """CachedNetEvaluator pads the miss list to a few batch sizes ({4,5,6,7} x 2^k >= 256): at most 25 % padding, never
    beyond the buffer, never below the number of misses."""
    seen = set()
    for cap in (400, 16384, 262144):
        for m in list(range(1, 3000)) + [4095, 4096, 4097, 100000, 262143, 262144]:
            if m > cap:
                continue
            b = ds._bucket(m, cap)
            assert m <= b <= cap
            if b < cap and m >= 256:
                assert b <= 1.25 * m + 1
                assert (b >> (b.bit_length() - 3)) in (4, 5, 6, 7) and b % (1 << (b.bit_length() - 3)) == 0
            if m <= 256 <= cap:
                assert b == 256
            seen.add(b)
    assert len(seen) < 60                                    # a few dozen shapes, whatever the miss counts


def test_net_evaluator_drops_graphs_when_weights_move():
    nets = importlib.import_module("alphazero-al_b200.nets")
    net = nets.C4Net()
    ev = ds.NetEvaluator(net)
    ev._check_weights()
    ev._graphs["captured"] = object()
    ev._check_weights()
    assert "captured" in ev._graphs                          # same storage: graphs stay
    net.load_state_dict(net.state_dict())                    # in-place update: graphs stay
    ev._check_weights()
    assert "captured" in ev._graphs
    net.half()                                               # new parameter storage: graphs must go
    ev._check_weights()
    assert not ev._graphs
