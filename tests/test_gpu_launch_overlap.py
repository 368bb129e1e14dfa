"""Launch policy (include/gnnfd_b200.h: gnnfd_set_launch_overlap, programmatic dependent launch): overlapped launches
give the stream-ordered results bit for bit - every kernel waits for its predecessor before its first global access."""
import pytest
import torch

from helpers import build_model, golden_graphs

pytestmark = pytest.mark.gpu


def _pin(monkeypatch, on):
    from gnn_fluid_dynamics_b200 import _lib, rollout, topology
    monkeypatch.setattr(topology, "choose_launch_overlap", lambda *a: None)
    monkeypatch.setattr(rollout, "choose_launch_overlap", lambda *a: None)
    _lib.set_launch_overlap(on)


@pytest.mark.parametrize("name", ["FvgnA", "MgnA", "ConservativeA"])
def test_training_step_bit_identical_with_overlapped_launches(name, monkeypatch):
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).train()
    _, graphs = golden_graphs(name, flip=True, n_cells=2000)
    runs = []
    for on in (False, True, False):
        _pin(monkeypatch, on)
        model.zero_grad(set_to_none=True)
        out = model([g.clone().to(dev) for g in graphs], mode="train")
        gn = model.normalizer.input([g.clone().to(dev) for g in graphs])
        loss = model.loss(out, gn)["total_log_loss"]
        loss.backward()
        runs.append([loss.detach().clone()] + [p.grad.clone() for p in model.parameters() if p.grad is not None])
    assert all(torch.equal(a, b) for a, b in zip(runs[0], runs[1]))
    assert all(torch.equal(a, b) for a, b in zip(runs[0], runs[2]))


@pytest.mark.parametrize("name", ["FvgnA", "MgnA", "FluxA"])
def test_rollout_bit_identical_with_overlapped_launches(name, monkeypatch):
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).eval()
    _, graphs = golden_graphs(name, flip=False, n_cells=400)
    outs = []
    for on in (False, True):
        _pin(monkeypatch, on)
        eng = RolloutEngine(model, [g.clone().to(dev) for g in graphs])      # captured with the pinned policy
        outs.append(torch.stack(eng.run(8, keep=True)))
    assert torch.equal(outs[0], outs[1])


def test_policy_choice():
    from gnn_fluid_dynamics_b200 import _lib
    _lib.choose_launch_overlap(3000, False)
    assert _lib._launch_overlap is True
    _lib.choose_launch_overlap(300000, False)
    assert _lib._launch_overlap is False
    _lib.choose_launch_overlap(300000, True)
    assert _lib._launch_overlap is True
    assert _lib.lib.gnnfd_set_launch_overlap(0) == 1 and _lib.lib.gnnfd_set_launch_overlap(0) == 0
    _lib._launch_overlap = None
