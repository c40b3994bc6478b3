"""ReferenceNetAdapter (device contract for an unmodified reference network) against the reference's own ``predict`` -
runs where /root/reference exists (this container), on the CPU; the GPU box has no reference checkout."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "environments")), reason="reference checkout not present")


def _load_ref_network(game):
    """Import src/environments/<game>/Network.py without the package __init__ (which needs the compiled env_cpp)."""
    saved = {k: sys.modules.get(k) for k in ("src", "src.environments", f"src.environments.{game}")}
    try:
        for name, path in (("src", f"{REF}/src"), ("src.environments", f"{REF}/src/environments"),
                           (f"src.environments.{game}", f"{REF}/src/environments/{game}")):
            pk = types.ModuleType(name)
            pk.__path__ = [path]
            sys.modules[name] = pk
        return importlib.import_module(f"src.environments.{game}.Network")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("game,shape,A", [("Connect4", (6, 7), 7), ("Othello", (8, 8), 65)])
def test_adapter_equals_reference_predict(game, shape, A):
    import torch
    ds = importlib.import_module("alphazero-al_b200.device_search")
    mod = _load_ref_network(game)
    torch.manual_seed(1)
    net = mod.CNN(lr=0, device="cpu")
    net.eval()
    with torch.no_grad():                                     # the heads start at zero: give them something to say
        for p in net.parameters():
            if p.abs().sum() == 0:
                p.normal_(0, 0.3)
    if game == "Othello":
        net.score_scale = 5.0
    rng = np.random.default_rng(2)
    B = 33
    own = rng.random((B, *shape)) < 0.3
    opp = (rng.random((B, *shape)) < 0.4) & ~own
    turn = rng.choice([-1.0, 1.0], size=B)
    planes = np.stack([own, opp, np.broadcast_to(turn[:, None, None], own.shape)], axis=1).astype(np.float32)
    mask = rng.random((B, A)) < 0.7
    mask[:, 0] = True
    want = net.predict(planes, mask)
    ad = ds.ReferenceNetAdapter(net, game)
    got = ad.predict_device(torch.from_numpy(planes), torch.from_numpy(mask.astype(np.uint8)))
    assert np.allclose(got[0].numpy(), want[0], rtol=0, atol=1e-6)
    assert np.allclose(got[1].numpy(), want[1], rtol=0, atol=1e-6)
    assert np.allclose(got[2].numpy().reshape(-1, 1), want[2], rtol=0, atol=1e-5)
    assert not ds.ReferenceNetAdapter.accepts(net)            # weights on the CPU: the wrapper keeps the host predict() path
    assert not ds.ReferenceNetAdapter.accepts(object())
