"""Fused finite-volume glue kernels (gnn_fluid_dynamics_b200/fvm_ops.py, csrc/glue.cu) against the oracle's plain tensor
restatement of the reference formulas (oracle/model.py: fvgn_integrator / fvgn_loss follow Fvgn.py:176-255,
normalisation.py:325-344, fvm.py:26-37, loss.py:55-60): forward values and autograd gradients, fp32, 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from fixtures import rel_l2
from helpers import golden_graphs

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _setup(n_cells=700, flip=True):
    from gnn_fluid_dynamics_b200.topology import get_topology
    _, graphs = golden_graphs("FvgnA", flip=flip, n_cells=n_cells, mesh_seed=5, feat_seed=6)
    gd = [g.to(dev()) for g in graphs]
    topo = get_topology(gd).validate()
    return graphs, gd, topo


def _ref_area(bn, f_area, vol, ei, dt, training):
    raw = (f_area * (torch.mean(dt) / ((vol[ei[0]] + vol[ei[1]]) / 2))).view(-1, 1)
    return F.batch_norm(raw, bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"], training=training,
                        momentum=0.1, eps=1e-5)


@pytest.mark.parametrize("training", [True, False])
def test_face_area_norm_forward_backward_and_running_stats(training):
    from gnn_fluid_dynamics_b200 import fvm_ops
    graphs, gd, topo = _setup()
    c, f, _ = graphs
    bn = torch.nn.BatchNorm1d(1)
    with torch.no_grad():
        bn.weight.fill_(1.3); bn.bias.fill_(-0.2); bn.running_mean.fill_(0.2); bn.running_var.fill_(1.5)
    ref = {k: v.detach().clone() for k, v in bn.state_dict().items()}
    ref["weight"].requires_grad_(True); ref["bias"].requires_grad_(True)
    dt = torch.tensor([0.01, 0.02, 0.015])
    out_ref = _ref_area(ref, f.area, c.volume, c.edge_index, dt, training)
    g = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(1))
    (out_ref * g).sum().backward()
    bn = bn.to(dev()).train(training)
    out = fvm_ops.face_area_norm(gd[1].area, gd[0].volume, topo.row, topo.col, dt.to(dev()), bn)
    (out * g.to(dev())).sum().backward()
    assert rel_l2(out, out_ref) < 1e-5
    assert rel_l2(bn.weight.grad, ref["weight"].grad) < 1e-5 and rel_l2(bn.bias.grad, ref["bias"].grad) < 1e-5
    assert rel_l2(bn.running_mean, ref["running_mean"]) < 1e-6 and rel_l2(bn.running_var, ref["running_var"]) < 1e-5
    assert int(bn.num_batches_tracked) == (1 if training else 0)


def test_fvm_integrate_and_divergence_vs_tensor_code():
    from gnn_fluid_dynamics_b200 import fvm_ops
    graphs, gd, topo = _setup()
    c, f, _ = graphs
    gen = torch.Generator().manual_seed(2)
    E, N = f.area.shape[0], c.x.shape[0]
    eo = torch.randn(E, 5, generator=gen).requires_grad_(True)
    area = (torch.rand(E, 1, generator=gen) + 0.5).requires_grad_(True)
    unv, cf = c.normal, f.face
    uv, p, fd = eo[:, :2], eo[:, 2:3], eo[:, 3:]
    uu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], -1)
    dot2 = lambda a, n: torch.cat([(a[:, 0:2] * n).sum(-1, keepdim=True), (a[:, 2:4] * n).sum(-1, keepdim=True)], -1)
    phi_a = sum(dot2(uu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
    phi_d = fd[cf[0]] + fd[cf[1]] + fd[cf[2]]
    phi_p = sum(p[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
    acc_ref = -phi_a - phi_p / 1.7 + phi_d
    div_ref = sum((uv[cf[j]] * unv[:, j, :]).sum(-1, keepdim=True) * area[cf[j]] for j in range(3))
    g1, g2 = torch.randn(N, 2, generator=gen), torch.randn(N, 1, generator=gen)
    ((acc_ref * g1).sum() + (div_ref * g2).sum()).backward()
    eo_d = eo.detach().to(dev()).requires_grad_(True)
    ar_d = area.detach().to(dev()).requires_grad_(True)
    cfd = fvm_ops.cell_faces(topo, gd[1].face)
    acc = fvm_ops.fvm_integrate(eo_d, ar_d, gd[0].normal, cfd, topo.row, topo.col, rho=1.7)
    div = fvm_ops.fvm_divergence(eo_d[:, :2], ar_d, gd[0].normal, cfd, topo.row, topo.col)
    ((acc * g1.to(dev())).sum() + (div * g2.to(dev())).sum()).backward()
    assert rel_l2(acc, acc_ref) < 1e-5 and rel_l2(div, div_ref) < 1e-5
    assert rel_l2(eo_d.grad, eo.grad) < 1e-5 and rel_l2(ar_d.grad, area.grad) < 1e-5


@pytest.mark.parametrize("masked", [False, True])
def test_masked_mse_vs_torch(masked):
    from gnn_fluid_dynamics_b200 import fvm_ops
    gen = torch.Generator().manual_seed(3)
    a = torch.randn(5001, 5, generator=gen).requires_grad_(True)
    b = torch.randn(5001, 3, generator=gen)
    mask = (torch.rand(5001, generator=gen) > 0.3) if masked else None
    av, bv = a[:, 1:3], b[:, :2]
    ref = F.mse_loss(av[mask], bv[mask]) if masked else F.mse_loss(av, bv)
    (ref * 1.7).backward()
    ad = a.detach().to(dev()).requires_grad_(True)
    out = fvm_ops.masked_mse(ad[:, 1:3], b.to(dev())[:, :2], None if mask is None else mask.to(dev()))
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    assert rel_l2(ad.grad, a.grad) < 1e-6
    out2 = fvm_ops.masked_mse(ad[:, 1:3].detach(), b.to(dev())[:, :2], None if mask is None else mask.to(dev()))
    assert torch.equal(out2, out.detach())                      # deterministic reduction


def test_state_advance_matches_update_features():
    """state_advance == rollout.py:340 + FvgnA.update_features + the z-scoring of the next step's inputs."""
    from gnn_fluid_dynamics_b200 import fvm_ops
    graphs, gd, topo = _setup(flip=False)
    c, f, _ = gd
    gen = torch.Generator().manual_seed(4)
    delta = torch.randn(c.x.shape[0], 2, generator=gen).to(dev())
    vel = c.x[:, :2] + delta
    dv = vel[c.edge_index[0]] - vel[c.edge_index[1]]
    mask = ((f.type == 2) | (f.type == 1)).reshape(-1)
    dv = torch.where(mask.unsqueeze(-1), f.y[:, 0:2], dv)
    x_raw, f_raw = c.x.clone(), f.x.clone()
    x_norm, f_norm = torch.zeros_like(x_raw), torch.zeros_like(f_raw)
    cs, fs = (0.1, 1.2, 0.15, 1.1), (0.3, 0.9, 0.25, 1.3)
    vout = torch.empty_like(vel)
    fvm_ops.state_advance(x_raw, delta, True, topo.row, topo.col, f_raw, mask, f.y, x_norm=x_norm, cell_stats=cs,
                          f_norm=f_norm, face_stats=fs, vel_out=vout)
    assert torch.equal(x_raw[:, :2], vel) and torch.equal(vout, vel) and torch.equal(f_raw[:, :2], dv)
    assert torch.equal(f_raw[:, 2:], f.x[:, 2:])
    assert rel_l2(x_norm[:, 0], (vel[:, 0] - cs[0]) / cs[1]) < 1e-6 and rel_l2(f_norm[:, 1], (dv[:, 1] - fs[2]) / fs[3]) < 1e-6


@pytest.mark.parametrize("name", ["FvgnA", "ConservativeA", "MgnC"])
def test_fused_normaliser_is_bit_identical_to_the_tensor_expressions(name):
    """Normalizer.input / output through gnnfd_affine_columns (one launch per tensor) == the per-column tensor expressions
    of normalisation.py:255-322, bit for bit, forward and inverse (z_score, mean_scale kinds; CPU path = the expressions)."""
    from helpers import build_model, golden_graphs
    model = build_model(name)
    _, graphs = golden_graphs(name, n_cells=500)
    ref = model.normalizer.input([g.clone() for g in graphs])                    # CPU: tensor expressions
    model.to(dev())
    got = model.normalizer.input([g.clone().to(dev()) for g in graphs])          # CUDA: fused kernel
    for a, b in zip(ref, got):
        for k in a.keys():
            if torch.is_tensor(a[k]) and a[k].is_floating_point():
                assert torch.equal(a[k], b[k].cpu()), k
    outs = [torch.randn(500, 5), torch.randn(700, 5), None]
    ref_o = model.cpu().normalizer.output([None if t is None else t.clone() for t in outs], inverse=True)
    model.to(dev())
    got_o = model.normalizer.output([None if t is None else t.clone().to(dev()) for t in outs], inverse=True)
    for a, b in zip(ref_o, got_o):
        if a is not None:
            assert torch.equal(a, b.cpu())


def test_flux_integrate_is_bit_identical_to_the_tensor_expression():
    """FluxA's integrator (Flux.py:166-206) and face_flux_to_cell_flux (fvm.py:96-156) as one kernel
    (gnnfd_flux_integrate) against the tensor expressions of models/Flux.py evaluated on the SAME device, bit for bit:
    the kernel rounds every product and sum separately, in the expression's order."""
    from gnn_fluid_dynamics_b200 import fvm_ops
    from gnn_fluid_dynamics_b200.models.Flux import face_flux_to_cell_flux
    graphs, gd, topo = _setup(n_cells=900, flip=True)
    c, f, _ = gd
    gen = torch.Generator().manual_seed(9)
    E, N = f.area.shape[0], c.x.shape[0]
    eo = torch.randn(E, 6, generator=gen).to(dev())
    coeff = torch.randn(E, 1, generator=gen).to(dev())
    area = torch.randn(E, 1, generator=gen).to(dev())
    unv, cf = c.normal, f.face
    rho = 1.3
    cell_flux = face_flux_to_cell_flux(eo[:, 3:4], cf, c.edge_index)
    uv, p_face, flux_d = eo[:, :2], eo[:, 2:3], eo[:, 4:6]
    phi_a = sum(uv[cf[j]] * cell_flux[:, j] * coeff[cf[j]] for j in range(3))
    phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
    phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
    ref = 1.0 * (-phi_a - phi_p / rho) + phi_d
    cfs = fvm_ops.cell_faces(topo, cf)
    acc, cfl = fvm_ops.flux_integrate(eo, coeff, area, unv, cfs, topo.row, topo.col, rho, want_cell_flux=True)
    assert torch.equal(acc, ref)
    assert torch.equal(cfl, cell_flux.squeeze(-1))
    # boundary faces are self-loops: the owner keeps +1, nobody gets -1
    boundary = (c.edge_index[0] == c.edge_index[1])
    assert bool(boundary.any())
    # de-normalised flux column of a wider matrix, cell flux only
    wide = torch.randn(E, 9, generator=gen).to(dev())
    got = fvm_ops.flux_integrate(wide, None, None, None, cfs, topo.row, topo.col, want_acc=False, want_cell_flux=True,
                                 flux_col=7, flux_scale=2.5, flux_shift=-0.75)
    assert torch.equal(got, face_flux_to_cell_flux(wide[:, 7:8] * 2.5 + (-0.75), cf, c.edge_index).squeeze(-1))


def test_flux_model_forward_with_the_kernel_equals_the_tensor_path():
    """FluxA.forward in evaluation mode (kernel integrator + kernel cell flux) == the same forward with the topology hidden
    from the integrator (tensor expressions): the face outputs and the cell flux bit for bit, the integrated change to
    1e-6 (with the topology hidden the face-area BatchNorm is ATen's, not fvm_ops.face_area_norm: last-bit differences)."""
    from helpers import build_model
    from gnn_fluid_dynamics_b200.models import Flux
    model = build_model("FluxA").to(dev()).eval()
    _, graphs = golden_graphs("FluxA", n_cells=800, mesh_seed=15, feat_seed=16)
    with torch.no_grad():
        a = model([g.clone().to(dev()) for g in graphs], mode="rollout")
        orig_topo, orig_ops = Flux.graph_topology, None
        Flux.graph_topology = lambda c_graph: None            # integrator -> tensor expressions
        try:
            b = model([g.clone().to(dev()) for g in graphs], mode="rollout")
        finally:
            Flux.graph_topology = orig_topo
    assert set(a) == set(b)
    for k in a:
        if k == "cell_flux":
            ref = Flux.face_flux_to_cell_flux(b["face_flux"], graphs[1].face.to(dev()), graphs[0].edge_index.to(dev()))
            assert torch.equal(a[k], ref.squeeze(-1))
        if k == "cell_velocity_change":
            assert rel_l2(a[k], b[k]) < 1e-6, rel_l2(a[k], b[k])
        else:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("width", [1, 2, 5, 8])
def test_gather3_forward_bitwise_and_backward_vs_index_put(width):
    """gnnfd_gather3: the three x[f_graph.face[j]] gathers as one launch (bit-identical to indexing) and its sort-free
    transpose against autograd's index_put backward (same sums, possibly another order: 1e-6)."""
    from gnn_fluid_dynamics_b200 import fvm_ops
    graphs, gd, topo = _setup(n_cells=1100, flip=True)
    c, f, _ = gd
    gen = torch.Generator().manual_seed(13)
    E, N = f.area.shape[0], c.x.shape[0]
    wide = torch.randn(E, width + 3, generator=gen).to(dev())
    t = wide[:, 1:1 + width].detach().requires_grad_(True)          # a column slice of a wider matrix
    t_ref = t.detach().clone().requires_grad_(True)
    cf = f.face
    out = fvm_ops.gather3(t, fvm_ops.cell_faces(topo, cf), topo.row, topo.col)
    ref = torch.stack([t_ref[cf[0]], t_ref[cf[1]], t_ref[cf[2]]])
    assert out.shape == (3, N, width) and torch.equal(out, ref)
    g = torch.randn(3, N, width, generator=gen).to(dev())
    (out * g).sum().backward()
    (ref * g).sum().backward()
    assert rel_l2(t.grad, t_ref.grad) < 1e-6
    # every face of a valid mesh is listed by its cells: the transpose touches every row
    assert bool((t.grad.abs().sum(1) > 0).all())
