"""GPU parity of the backward kernels (BASELINE.json north_star (e)): unit kernels against plain fp64/fp32
torch restatements, the whole training step (loss + every parameter gradient) against the CPU oracle's
autograd and against the reference-generated golden fixture tests/golden/train_FvgnA.npz.

Tolerances: loss 1e-4 relative; gradients rel-L2 <= 1e-3 per parameter tensor (same bar north_star states
for processor outputs: dgrad runs split-bf16 (~1e-5), wgrad runs single-pass TF32 with round-to-nearest
operands (~3e-4), fp32 accumulation everywhere)."""
import pytest
import torch

from oracle import model as omodel
from fixtures import default_stats, rel_l2
from helpers import LOSS_W, build_model, golden_graphs, load_golden

pytestmark = pytest.mark.gpu
GRAD_TOL = 1e-3     # whole-step parameter gradients; measured worst case ~1e-4 (split-bf16 dgrad and wgrad)


def dev():
    return torch.device("cuda:0")


def i32(t):
    return t.to(torch.int32).to(dev())


# ------------------------------------------------------------------------------------ unit kernels
@pytest.mark.parametrize("rows", [1, 31, 32, 33, 1000, 40000])
def test_wgrad_dense_vs_torch(rows):
    from gnn_fluid_dynamics_b200 import ops
    g = torch.Generator().manual_seed(rows)
    a, b = torch.randn(rows, 128, generator=g), torch.randn(rows, 128, generator=g)
    ref = a.double().t() @ torch.nn.functional.silu(b.double())
    out = torch.empty(128, 128, device=dev())
    cs = torch.empty(128, device=dev())
    ops.wgrad(ops.Seg(a.to(dev())), [ops.Seg(b.to(dev()))], rows, out, b_act=1, colsum=cs)
    assert rel_l2(out, ref.float()) < 5e-5, rel_l2(out, ref.float())          # split-bf16 (default): ~1e-5
    assert rel_l2(cs, a.double().sum(0).float()) < 1e-5
    ops.wgrad(ops.Seg(a.to(dev())), [ops.Seg(b.to(dev()))], rows, out, b_act=1, single_pass=True)
    assert rel_l2(out, ref.float()) < GRAD_TOL, rel_l2(out, ref.float())     # single-pass TF32


def test_wgrad_gathered_concat_and_narrow_vs_torch():
    """dW1 of a Face_Block (B = [e | x[row] | x[col]], 384 columns), of a Cell_Block (B = [x | mean3(vsum)]),
    of an encoder (B 10 columns) and dW3 of a decoder head (narrow side transposed)."""
    from gnn_fluid_dynamics_b200 import ops, _lib
    g = torch.Generator().manual_seed(11)
    N, E, V = 700, 1111, 400
    x, e, vs = torch.randn(N, 128, generator=g), torch.randn(E, 128, generator=g), torch.randn(V, 64, generator=g)
    row, col = torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
    vf = [torch.randint(0, V, (N,), generator=g) for _ in range(3)]
    da_e, da_n = torch.randn(E, 128, generator=g), torch.randn(N, 128, generator=g)
    xd, ed, vsd = x.to(dev()), e.to(dev()), vs.to(dev())
    # edge layer 1
    ref = da_e.double().t() @ torch.cat([e, x[row], x[col]], 1).double()
    out = torch.empty(128, 384, device=dev())
    cs = torch.empty(128, device=dev())
    ops.wgrad(ops.Seg(da_e.to(dev())), [ops.Seg(ed), ops.Seg(xd, _lib.SEG_GATHER, (i32(row),)),
                                        ops.Seg(xd, _lib.SEG_GATHER, (i32(col),))], E, out, colsum=cs)
    assert rel_l2(out, ref.float()) < 5e-5, rel_l2(out, ref.float())
    assert rel_l2(cs, da_e.double().sum(0).float()) < 1e-5
    # node layer 1
    agg = (vs[vf[0]] + vs[vf[1]] + vs[vf[2]]) / 3.0
    ref = da_n.double().t() @ torch.cat([x, agg], 1).double()
    out = torch.empty(128, 192, device=dev())
    ops.wgrad(ops.Seg(da_n.to(dev())), [ops.Seg(xd), ops.Seg(vsd, _lib.SEG_MEAN3, tuple(i32(t) for t in vf))], N, out)
    assert rel_l2(out, ref.float()) < 5e-5, rel_l2(out, ref.float())
    # encoder layer 1 (10 input columns) and decoder layer 3 (5 outputs, transposed store, colsum of B)
    f = torch.randn(E, 10, generator=g)
    out = torch.empty(128, 10, device=dev())
    ops.wgrad(ops.Seg(da_e.to(dev())), [ops.Seg(f.to(dev()))], E, out)
    assert rel_l2(out, (da_e.double().t() @ f.double()).float()) < GRAD_TOL
    dy = torch.randn(E, 5, generator=g)
    out = torch.empty(5, 128, device=dev())
    cs = torch.empty(5, device=dev())
    ops.wgrad(ops.Seg(ed), [ops.Seg(dy.to(dev()))], E, out, a_act=1, transpose_out=True, colsum=cs, colsum_of_b=True)
    assert rel_l2(out, (dy.double().t() @ torch.nn.functional.silu(e.double())).float()) < GRAD_TOL
    assert rel_l2(cs, dy.double().sum(0).float()) < 1e-5


@pytest.mark.parametrize("rows", [1, 7, 4096, 33333])
def test_ln_backward_vs_autograd(rows):
    from gnn_fluid_dynamics_b200 import ops
    g = torch.Generator().manual_seed(rows)
    y = torch.randn(rows, 128, generator=g, dtype=torch.float64, requires_grad=True)
    w = (1 + 0.1 * torch.randn(128, generator=g, dtype=torch.float64)).requires_grad_()
    b = torch.zeros(128, dtype=torch.float64, requires_grad=True)
    go = torch.randn(rows, 128, generator=g, dtype=torch.float64)
    out = torch.nn.functional.layer_norm(y, (128,), w, b, 1e-5)
    out.backward(go)
    mean, var = y.mean(1, keepdim=True), y.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    xhat = ((y - mean) * rstd).detach()
    dy, sums = ops.ln_backward(go.float().to(dev()), xhat.float().to(dev()), rstd.detach().float().reshape(-1).to(dev()),
                               w.detach().float().to(dev()))
    assert rel_l2(dy, y.grad.float()) < 1e-5
    assert rel_l2(sums[0], w.grad.float()) < 1e-5 and rel_l2(sums[1], b.grad.float()) < 1e-5
    assert rel_l2(sums[2], y.grad.sum(0).float()) < 1e-4


def test_linear_tc_dgrad_with_activation_derivative():
    from gnn_fluid_dynamics_b200 import ops, _lib
    g = torch.Generator().manual_seed(5)
    R = 777
    da, a_pre, res = torch.randn(R, 128, generator=g), torch.randn(R, 128, generator=g), torch.randn(R, 128, generator=g)
    W = torch.randn(128, 384, generator=g) * 0.1          # forward weight [out=128, in=384]
    ap = a_pre.double().requires_grad_()
    torch.nn.functional.silu(ap).backward(torch.ones_like(ap))
    Wd = W.to(dev())
    packs = {}
    # dIn segment 1 = dA . W[:, 128:256], accumulated onto a residual
    out = ops.linear_tc(ops.Seg(da.to(dev())), R, Wd[:, 128:], 1, 384, 128, 128, packs, ("w1t", 128), _lib.PREC_BF16X3,
                        residual=res.to(dev()))
    assert rel_l2(out, (res.double() + da.double() @ W[:, 128:256].double()).float()) < 1e-4
    # narrow block (64 valid output columns) with the activation derivative fused
    out = ops.linear_tc(ops.Seg(da.to(dev())), R, Wd[:, 320:], 1, 384, 64, 128, packs, ("w1t", 320), _lib.PREC_BF16X3,
                        mul=a_pre.to(dev()), mul_mode=1)
    ref = (da.double() @ W[:, 320:384].double()) * ap.grad[:, :64]
    assert rel_l2(out[:, :64], ref.float()) < 1e-4
    assert float(out[:, 64:].abs().max()) == 0.0


def test_transposed_gathers_vs_index_add():
    from gnn_fluid_dynamics_b200 import ops
    from gnn_fluid_dynamics_b200.mesh import make_mesh
    m = make_mesh(3000, "cylinder", seed=4)
    N, E, V = m.n_cells, m.n_faces, m.n_vertices
    g = torch.Generator().manual_seed(2)
    row, col = torch.from_numpy(m.cell_edge_index[0]), torch.from_numpy(m.cell_edge_index[1])
    t1, t2, base = torch.randn(E, 128, generator=g), torch.randn(E, 128, generator=g), torch.randn(N, 128, generator=g)
    off, perm = ops.csr_build(i32(torch.cat([row, col])), N)
    out = ops.segment_sum3(t1.to(dev()), t2.to(dev()), None, (0, 0, 0), 128, 1.0, E, off, perm, N, base=base.to(dev()))
    ref = base.double().index_add(0, row, t1.double()).index_add(0, col, t2.double())
    assert rel_l2(out, ref.float()) < 1e-6
    vf = [torch.from_numpy(m.cells[:, j].copy()).long() for j in range(3)] if hasattr(m, "cells") else None
    if vf is None:
        vf = [torch.randint(0, V, (N,), generator=g) for _ in range(3)]
    t3 = torch.randn(N, 128, generator=g)
    off, perm = ops.csr_build(i32(torch.cat(vf)), V)
    out = ops.segment_sum3(t3.to(dev()), None, None, (0, 0, 0), 64, 1.0, N, off, perm, V, scale=1.0 / 3.0)
    ref = torch.zeros(V, 64, dtype=torch.float64)
    for j in range(3):
        ref.index_add_(0, vf[j], t3[:, :64].double())
    assert rel_l2(out, (ref / 3).float()) < 1e-6
    v0, v1 = torch.from_numpy(m.vertex_edge_index[0]), torch.from_numpy(m.vertex_edge_index[1])
    dv, de = torch.randn(V, 64, generator=g), torch.randn(E, 128, generator=g)
    out = ops.gather_pair_add(dv.to(dev()), i32(v0), i32(v1), 1.0, True, E, base=de.to(dev()))
    assert torch.equal(out.cpu(), de + torch.cat([dv[v0], dv[v1]], 1))
    dn = torch.randn(N, 128, generator=g)
    out = ops.gather_pair_add(dn.to(dev()), i32(col), i32(row), -1.0, False, E)
    assert rel_l2(out, dn[col] - dn[row]) < 1e-6


# ---------------------------------------------------------------------------------- whole training step
def _oracle_step(name, model, graphs):
    params = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    out, _ = omodel.model_forward(name, params, default_stats(), [g.clone() for g in graphs], 15, mode="train", training=True)
    graphs_n = omodel.normalise_inputs(name, default_stats(), [g.clone() for g in graphs])
    if name == "FvgnA":
        loss = omodel.fvgn_loss(params, out, graphs_n, LOSS_W, training=True)["total_log_loss"]
    else:
        c = graphs_n[0]
        mse = lambda a, b: torch.mean((a - b) ** 2)
        total = LOSS_W["cell_velocity_change"] * mse(out["cell_velocity_change"], c.y[:, 0:2]) \
            + LOSS_W["cell_pressure"] * mse(out["cell_pressure"], c.y[:, 2:3])
        loss = torch.mean(torch.log(total))
    loss.backward()
    return float(loss), {k: p.grad for k, p in params.items() if p.grad is not None}


@pytest.mark.parametrize("name,n_cells", [("FvgnA", 160), ("MgnA", 160), ("FvgnA", 3000), ("FvgnA", 20000)])
def test_training_step_gradients_vs_oracle(name, n_cells):
    model = build_model(name).train()
    _, graphs = golden_graphs(name, flip=True, n_cells=n_cells)
    ref_loss, ref_grads = _oracle_step(name, model, graphs)
    model.to(dev())
    out = model([g.clone().to(dev()) for g in graphs], mode="train")
    gn = model.normalizer.input([g.clone().to(dev()) for g in graphs])
    loss = model.loss(out, gn)["total_log_loss"]
    loss.backward()
    assert abs(float(loss) - ref_loss) < 1e-4 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        if k not in ref_grads:
            continue
        assert p.grad is not None, k
        err = rel_l2(p.grad, ref_grads[k])
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < GRAD_TOL, worst


def test_training_step_matches_reference_golden():
    """Loss values and gradients of the reference's own FvgnA training step (tests/golden/train_FvgnA.npz)."""
    gold = load_golden("train_FvgnA.npz")
    model = build_model("FvgnA").to(dev()).train()
    _, graphs = golden_graphs("FvgnA", flip=True)
    out = model([g.clone().to(dev()) for g in graphs], mode="train")
    gn = model.normalizer.input([g.clone().to(dev()) for g in graphs])
    losses = model.loss(out, gn)
    for k, v in losses.items():
        assert abs(float(v) - float(gold[f"loss_{k}"][0])) < 1e-4 * max(1.0, abs(float(gold[f"loss_{k}"][0]))), k
    losses["total_log_loss"].backward()
    grads = dict(model.named_parameters())
    for n, ref_norm in zip([str(n) for n in gold["grad_names"]], gold["grad_norms"]):
        if n in grads and grads[n].grad is not None:
            assert abs(float(grads[n].grad.double().norm()) - ref_norm) <= 2e-3 * max(ref_norm, 1e-6) + 1e-9, n
    for k in gold:
        if k.startswith(("grad_processer", "grad_decoder", "grad_encoder")):
            assert rel_l2(grads[k[5:]].grad, torch.from_numpy(gold[k])) < GRAD_TOL, k


@pytest.mark.parametrize("name", ["FvgnA", "VertPotA", "StreamFuncA", "FluxA", "ConservativeA", "MgnA"])
def test_training_step_matches_reference_golden_families(name):
    """The reference's own training step (forward 'train' + model.loss + backward; tests/golden/make_golden.py
    --train-only) for BASELINE.json config 5's families (VertPotA, StreamFuncA) and FluxA / ConservativeA / MgnA:
    loss values, the norm of every live parameter gradient and six full gradient tensors."""
    gold = load_golden(f"train_{name}.npz")
    model = build_model(name).to(dev()).train()
    _, graphs = golden_graphs(name, flip=True)
    batch = [g.clone().to(dev()) for g in graphs]
    out = model(batch, mode="train")            # normalises the batch in place, like the reference (train.py:253)
    losses = model.loss(out, batch)
    for k, v in losses.items():
        ref = float(gold[f"loss_{k}"][0])
        assert abs(float(v) - ref) < 2e-4 * max(1.0, abs(ref)) or abs(ref) < 1e-12, (k, float(v), ref)
    losses["total_log_loss"].backward()
    grads = dict(model.named_parameters())
    checked = 0
    for n, ref_norm in zip([str(n) for n in gold["grad_names"]], gold["grad_norms"]):
        assert grads[n].grad is not None, n
        assert abs(float(grads[n].grad.double().norm()) - ref_norm) <= 2e-3 * max(ref_norm, 1e-6) + 1e-9, n
        checked += 1
    assert checked > 200
    for k in gold:
        if k.startswith("grad_") and k not in ("grad_names", "grad_norms"):
            assert rel_l2(grads[k[5:]].grad, torch.from_numpy(gold[k])) < GRAD_TOL, k


def test_training_gradients_are_bitwise_reproducible():
    model = build_model("FvgnA").to(dev()).train()
    _, graphs = golden_graphs("FvgnA", flip=True, n_cells=2000)
    runs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out = model([g.clone().to(dev()) for g in graphs], mode="train")
        gn = model.normalizer.input([g.clone().to(dev()) for g in graphs])
        model.loss(out, gn)["total_log_loss"].backward()
        runs.append([p.grad.clone() for p in model.parameters() if p.grad is not None])
    assert all(torch.equal(a, b) for a, b in zip(*runs))


def test_fused_backward_call_equals_stepwise_schedule():
    """gnnfd_mlp_backward (one C call, dgrad chain fused into one 3-layer pass) against the Python-scheduled
    chain of single-Linear launches: same arithmetic up to accumulation order."""
    from gnn_fluid_dynamics_b200 import ops, _lib, training
    from test_gpu_parity import _rand_mlp, _to_weights
    g = torch.Generator().manual_seed(21)
    N, E = 500, 900
    x, e = torch.randn(N, 128, generator=g).to(dev()), torch.randn(E, 128, generator=g).to(dev())
    row, col = i32(torch.randint(0, N, (E,), generator=g)), i32(torch.randint(0, N, (E,), generator=g))
    segs = [ops.Seg(e), ops.Seg(x, _lib.SEG_GATHER, (row,)), ops.Seg(x, _lib.SEG_GATHER, (col,))]
    w = _to_weights(_rand_mlp(384, 128, True, seed=3), _lib.ACT_SILU)
    _, _, st = ops.mlp_forward(segs, w, E, _lib.PREC_BF16X3, residual=e, want_raw=False, want_sum=True, stash=True)
    go, res = torch.randn(E, 128, generator=g).to(dev()), torch.randn(E, 128, generator=g).to(dev())
    ws = ops.mlp_backward_workspace(E, dev())
    ga, da = training.mlp_backward(w, st, segs, E, go, _lib.PREC_BF16X3, [{"residual": res}, {}, None], ws)
    gb, db = training.mlp_backward_stepwise(w, st, segs, E, go, _lib.PREC_BF16X3, [{"residual": res}, {}, None])
    for p, q in zip(ga, gb):
        assert (p is None) == (q is None)
        if p is not None:
            assert rel_l2(p, q) < 2e-5, rel_l2(p, q)
    assert rel_l2(da[0], db[0]) < 2e-5 and rel_l2(da[1], db[1]) < 2e-5 and da[2] is None and db[2] is None


def test_backward_entry_points_handle_empty_and_tiny_inputs():
    """rows = 0 (an empty mesh partition) and rows = 1: statuses, zero gradients, no launch faults."""
    from gnn_fluid_dynamics_b200 import ops, _lib, training
    from test_gpu_parity import _rand_mlp, _to_weights
    w = _to_weights(_rand_mlp(128, 128, True, seed=4), _lib.ACT_SILU)
    for rows in (0, 1):
        x = torch.randn(rows, 128, device=dev())
        segs = [ops.Seg(x)]
        _, _, st = ops.mlp_forward(segs, w, rows, _lib.PREC_BF16X3, stash=True)
        ws = ops.mlp_backward_workspace(max(rows, 1), dev())
        grads, dins = training.mlp_backward(w, st, segs, rows, torch.randn(rows, 128, device=dev()), _lib.PREC_BF16X3, [{}], ws)
        torch.cuda.synchronize()
        assert all(torch.isfinite(g).all() for g in grads if g is not None)
        if rows == 0:
            assert all(float(g.abs().sum()) == 0.0 for g in grads if g is not None)
        assert dins[0].shape == (rows, 128)
    out = torch.empty(128, 128, device=dev())
    ops.wgrad(ops.Seg(torch.empty(0, 128, device=dev())), [ops.Seg(torch.empty(0, 128, device=dev()))], 0, out)
    assert float(out.abs().sum()) == 0.0


def test_vertpot_processor_gradients_vs_oracle():
    """VertPot family (FVGN blocks + Vertex_Block + two decoder heads): gradients of a random linear functional of
    both heads w.r.t. every live parameter against the oracle's autograd."""
    import oracle
    from gnn_fluid_dynamics_b200.topology import get_topology
    name = "VertPotA"
    model = build_model(name).train()
    _, graphs = golden_graphs(name, n_cells=400)
    graphs = model.normalizer.input([g.clone() for g in graphs])
    c, f, v = graphs
    params = {k: p.detach().clone().requires_grad_(p.is_floating_point()) for k, p in model.state_dict().items()}
    topo_cpu = {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face, "n_vertices": v.num_nodes}
    ref = oracle.processor_fwd("vertpot", params, c.x, f.x, topo_cpu, 15)
    edge_ref, vert_ref = ref["dec"]
    gen = torch.Generator().manual_seed(3)
    r1, r2 = torch.randn(edge_ref.shape, generator=gen), torch.randn(vert_ref.shape, generator=gen)
    ((edge_ref * r1).sum() + (vert_ref * r2).sum()).backward()
    model.to(dev())
    gd = [g.to(dev()) for g in graphs]
    _, _, _, edge_out, vert_out = model.encode_process_decode(gd[0].x, gd[1].x, get_topology(gd).validate())
    assert rel_l2(edge_out, edge_ref.detach()) < 2e-3 and rel_l2(vert_out, vert_ref.detach()) < 2e-3
    ((edge_out * r1.to(dev())).sum() + (vert_out * r2.to(dev())).sum()).backward()
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        g_ref = params[k].grad
        if g_ref is None or float(g_ref.abs().max()) == 0.0:
            continue                 # the reference's unused duplicate face_block / cell_block parameters
        assert p.grad is not None, k
        err = rel_l2(p.grad, g_ref)
        worst = max(worst, (k, err), key=lambda t: t[1])
    assert worst[1] < GRAD_TOL, worst


@pytest.mark.parametrize("name", ["ConservativeA", "ConservativeD", "ConservativeE", "ConservativeF", "ConservativeG",
                                  "ConservativeI", "ConservativeH"])
def test_generic_autograd_path_gradients_vs_oracle(name):
    """Families without a hand-scheduled backward train through the per-op autograd wrappers (autograd_ops.py):
    gradients of a random linear functional of the decoder output w.r.t. every live parameter vs the oracle."""
    import oracle
    from gnn_fluid_dynamics_b200.topology import get_topology
    model = build_model(name).train()
    _, graphs = golden_graphs(name, n_cells=300)
    graphs = model.normalizer.input([g.clone() for g in graphs])
    c, f, v = graphs
    params = {k: p.detach().clone().requires_grad_(p.is_floating_point()) for k, p in model.state_dict().items()}
    topo_cpu = {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face, "n_vertices": v.num_nodes}
    dual = name in ("ConservativeA", "ConservativeD", "ConservativeH")
    bc = ((f.type == 2) | (f.type == 1)).reshape(-1) if name == "ConservativeI" else None
    ref = oracle.processor_fwd(oracle.family_of(name), params, c.x, f.x_symm if dual else f.x, topo_cpu, 15,
                               f_x_asym=f.x_asym if dual else None, bc_mask=bc)
    r = torch.randn(ref["dec"].shape, generator=torch.Generator().manual_seed(4))
    (ref["dec"] * r).sum().backward()
    model.to(dev())
    gd = [g.to(dev()) for g in graphs]
    topo = get_topology(gd, need_cell_csr=True, two_hop=(not dual) or name == "ConservativeH").validate()
    if dual:
        out = model.encode_process_decode(gd[0].x, gd[1].x_symm, gd[1].x_asym, topo)[2]
    elif name == "ConservativeI":
        keep = (~bc).float().unsqueeze(1).expand(-1, 128).contiguous().to(dev())
        out = model.encode_process_decode(gd[0].x, gd[1].x, topo, e_keep=keep)[2]
    else:
        out = model.encode_process_decode(gd[0].x, gd[1].x, topo)[2]
    assert rel_l2(out, ref["dec"].detach()) < 2e-3
    (out * r.to(dev())).sum().backward()
    worst, checked = ("", 0.0), 0
    for k, p in model.named_parameters():
        g_ref = params[k].grad
        if g_ref is None or float(g_ref.abs().max()) == 0.0:
            continue
        assert p.grad is not None, k
        worst = max(worst, (k, rel_l2(p.grad, g_ref)), key=lambda t: t[1])
        checked += 1
    assert checked > 100 and worst[1] < GRAD_TOL, (checked, worst)
