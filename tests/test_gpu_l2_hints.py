"""L2 eviction-priority hints (include/gnnfd_b200.h: gnnfd_set_l2_hints): the hint is an operand of the same loads /
stores / TMA copies, so every mask gives the unhinted results bit for bit."""
import os

import pytest
import torch

from helpers import build_model, golden_graphs

pytestmark = pytest.mark.gpu

ALL_HINTS = 31


@pytest.fixture
def hints():
    from gnn_fluid_dynamics_b200 import _lib
    if os.environ.get("GNNFD_L2_HINTS") is not None:
        pytest.skip("GNNFD_L2_HINTS pins the mask for the process")
    prev = _lib.lib.gnnfd_set_l2_hints(-1)
    yield lambda mask: _lib.lib.gnnfd_set_l2_hints(mask)
    _lib.lib.gnnfd_set_l2_hints(-1)
    assert prev == _lib.lib.gnnfd_set_l2_hints(-1)      # the library's default is back in force


@pytest.mark.parametrize("name", ["FvgnA", "MgnA"])
def test_training_step_bit_identical_under_every_mask(name, hints):
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).train()
    _, graphs = golden_graphs(name, flip=True, n_cells=2000)
    runs = []
    for mask in (0, ALL_HINTS, 5, 0):
        hints(mask)
        model.zero_grad(set_to_none=True)
        out = model([g.clone().to(dev) for g in graphs], mode="train")
        gn = model.normalizer.input([g.clone().to(dev) for g in graphs])
        loss = model.loss(out, gn)["total_log_loss"]
        loss.backward()
        runs.append([loss.detach().clone()] + [p.grad.clone() for p in model.parameters() if p.grad is not None])
    for other in runs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(runs[0], other))


@pytest.mark.parametrize("name", ["FvgnA", "FluxA", "ConservativeA"])
def test_rollout_bit_identical_under_every_mask(name, hints):
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).eval()
    _, graphs = golden_graphs(name, flip=False, n_cells=400)
    outs = []
    for mask in (0, ALL_HINTS, 10):
        hints(mask)
        eng = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=False)
        outs.append(torch.stack(eng.run(6, keep=True)))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_mask_setter_returns_previous(hints):
    assert hints(7) >= 0
    assert hints(3) == 7
    assert hints(-1) == 3
