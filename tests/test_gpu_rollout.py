"""100-step autoregressive rollout: CUDA path vs the CPU oracle within 1e-2 rel-L2 (BASELINE.json north_star),
CUDA-graph replay bit-identical to eager stepping."""
import pytest
import torch

from oracle import model as omodel
from gnn_fluid_dynamics_b200.testing import default_stats, rel_l2
from helpers import build_model, golden_graphs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["FvgnA", "MgnA"])
def test_rollout_100_steps_vs_oracle(name):
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    dev = torch.device("cuda:0")
    model = build_model(name).eval()
    _, graphs = golden_graphs(name, n_cells=400, mesh_seed=31, feat_seed=32)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref_graphs = [g.clone() for g in graphs]
    with torch.no_grad():
        for _ in range(100):
            ref = omodel.rollout_step(name, sd, default_stats(), ref_graphs, 15)
    assert torch.isfinite(ref).all()
    model.to(dev)
    eng = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=True)
    vel = eng.run(100, keep=True)[-1]
    err = rel_l2(vel, ref)
    assert err < 1e-2, err
    eager = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=False)
    vel_e = eager.run(100, keep=True)[-1]
    assert torch.equal(vel, vel_e)            # graph replay == eager stepping, bit for bit


@pytest.mark.parametrize("name", ["FluxA", "ConservativeA", "ConservativeD"])
def test_rollout_engine_other_families_graph_equals_eager(name):
    """BASELINE.json config 3 families (Conservative / Flux face-flux models): 20 autoregressive steps, CUDA-graph
    replay bit-identical to eager stepping, finite state."""
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).eval()
    _, graphs = golden_graphs(name, n_cells=500, mesh_seed=51, feat_seed=52)
    cons = name.startswith("Conservative")
    kw = dict(need_cell_csr=cons, two_hop=not cons)
    a = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=True, **kw)
    b = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=False, **kw)
    va = a.run(20, keep=True)
    vb = b.run(20, keep=True)
    assert all(torch.equal(p, q) for p, q in zip(va, vb))
    assert torch.isfinite(va[-1]).all()
    assert not torch.equal(va[0], va[-1])
