"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the reference-made
golden fixtures.  Tolerances: integer work bit-exact; fp32 path 2e-5 rel-L2 (summation order only);
tensor-core precisions <= 1e-3 rel-L2 per processor output (BASELINE.json: north_star)."""
import numpy as np
import pytest
import torch

import oracle
from fixtures import rel_l2
from helpers import ALL_MODELS, LOSS_MODELS, build_model, golden_graphs, load_golden

pytestmark = pytest.mark.gpu

# fp16x2 (fp16 weights, split activations) and bf16x1 are built but measured at 4.6e-3 / >1e-2 on the
# FvgnA forward: they fail the bar and are therefore not parity-gated defaults.
TOL = {"f32": 2e-5, "bf16x3": 1e-3, "fp16x3": 1e-3}


def precisions():
    from gnn_fluid_dynamics_b200.precisions import available
    return [p for p in available() if p in TOL]


def dev():
    return torch.device("cuda:0")


# ---------------------------------------------------------------------------------- CSR (integer)
@pytest.mark.parametrize("n,rows", [(0, 5), (1, 1), (1000, 37), (4096, 4096), (100000, 3), (50000, 70001)])
def test_csr_build_is_stable_argsort(n, rows):
    from gnn_fluid_dynamics_b200 import ops
    g = torch.Generator().manual_seed(n + rows)
    idx = torch.randint(0, rows, (n,), generator=g, dtype=torch.int64)
    off, perm = ops.csr_build(ops.index_narrow(idx.to(dev()), rows), rows)
    ref_off, ref_perm = oracle.csr_build(idx.numpy(), rows)
    assert np.array_equal(off.cpu().numpy(), ref_off)
    assert np.array_equal(perm.cpu().numpy(), ref_perm)
    if n:
        assert torch.equal(perm.cpu().long(), torch.sort(idx, stable=True).indices)


def test_csr_on_mesh_index_vectors_and_range_check():
    from gnn_fluid_dynamics_b200.mesh import make_mesh
    from gnn_fluid_dynamics_b200 import ops
    m = make_mesh(20000, "cylinder", seed=1)
    for vec, rows in [(np.concatenate([m.vertex_edge_index[0], m.vertex_edge_index[1]]), m.n_vertices),
                      (np.concatenate([m.cell_edge_index[1], m.cell_edge_index[0]]), m.n_cells)]:
        off, perm = ops.csr_build(ops.index_narrow(torch.from_numpy(vec).to(dev()), rows), rows)
        ref_off, ref_perm = oracle.csr_build(vec, rows)
        assert np.array_equal(off.cpu().numpy(), ref_off) and np.array_equal(perm.cpu().numpy(), ref_perm)
    bad = ops.index_narrow(torch.tensor([0, 5, 9], device=dev()), 9)
    assert int(bad._gnnfd_range_flag.item()) == 1


# -------------------------------------------------------------------- segment sum (bit-exact order)
@pytest.mark.parametrize("mode", ["halves", "signed", "full"])
def test_segment_sum_matches_sequential_scatter_add(mode):
    from gnn_fluid_dynamics_b200 import ops
    from gnn_fluid_dynamics_b200.mesh import make_mesh
    m = make_mesh(3000, "airfoil", seed=2)
    E = m.n_faces
    e = torch.randn(E, 128, generator=torch.Generator().manual_seed(0))
    if mode == "signed":
        idx = np.concatenate([m.cell_edge_index[1], m.cell_edge_index[0]]); rows = m.n_cells
        src = np.concatenate([e.numpy(), -e.numpy()]); args = (0, 0, 128, -1.0)
    elif mode == "halves":
        idx = np.concatenate([m.vertex_edge_index[0], m.vertex_edge_index[1]]); rows = m.n_vertices
        src = np.concatenate([e.numpy()[:, :64], e.numpy()[:, 64:]]); args = (0, 64, 64, 1.0)
    else:
        idx = np.concatenate([m.vertex_edge_index[0], m.vertex_edge_index[1]]); rows = m.n_cells
        src = np.concatenate([e.numpy(), e.numpy()]); args = (0, 0, 128, 1.0)
    ref = oracle.scatter_add_loop(src, idx, rows)
    off, perm = ops.csr_build(ops.index_narrow(torch.from_numpy(idx).to(dev()), rows), rows)
    ed = e.to(dev())
    out = ops.segment_sum(ed, ed, args[0], args[1], args[2], args[3], off, perm, rows)
    assert np.array_equal(out.cpu().numpy(), ref)      # same order of additions -> bit-identical


# -------------------------------------------------------------------------------- fused MLP block
def _rand_mlp(k_in, n_out, ln, bias=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, b: (torch.rand(*s, generator=g) * 2 - 1) * b
    p = dict(w1=u(128, k_in, b=k_in ** -0.5), b1=u(128, b=k_in ** -0.5) if bias else None,
             w2=u(128, 128, b=128 ** -0.5), b2=u(128, b=128 ** -0.5) if bias else None,
             w3=u(n_out, 128, b=128 ** -0.5), b3=u(n_out, b=128 ** -0.5) if bias else None,
             ln_w=1 + u(n_out, b=0.1) if ln else None, ln_b=u(n_out, b=0.1) if ln else None)
    return p


def _to_weights(p, act):
    from gnn_fluid_dynamics_b200.ops import MLPWeights
    d = lambda t: None if t is None else t.to(dev()).contiguous()
    return MLPWeights(w1=d(p["w1"]), b1=d(p["b1"]), w2=d(p["w2"]), b2=d(p["b2"]), w3=d(p["w3"]),
                      b3=d(p["b3"]), ln_w=d(p["ln_w"]), ln_b=d(p["ln_b"]), has_ln=p["ln_w"] is not None,
                      act=act)


@pytest.mark.parametrize("rows", [1, 63, 64, 65, 1000])
@pytest.mark.parametrize("k_in,n_out,ln,act,bias", [
    (128, 128, True, "silu", True), (10, 128, True, "silu", True), (2, 128, True, "silu", True),
    (8, 128, True, "silu", True), (4, 128, False, "tanh", False), (128, 5, False, "silu", True),
    (128, 3, False, "silu", True), (128, 1, False, "silu", True), (128, 6, False, "silu", True)])
def test_mlp_rows_vs_oracle(rows, k_in, n_out, ln, act, bias):
    from gnn_fluid_dynamics_b200 import ops, _lib
    for prec in precisions():
        p = _rand_mlp(k_in, n_out, ln, bias, seed=rows + k_in)
        x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(1))
        ref = oracle.mlp3(x, p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"], act=act)
        w = _to_weights(p, _lib.ACT_SILU if act == "silu" else _lib.ACT_TANH)
        out, _ = ops.mlp_forward([ops.Seg(x.to(dev()))], w, rows, _lib.PRECISIONS[prec])
        assert rel_l2(out, ref) < TOL[prec], (prec, rel_l2(out, ref))


def test_fused_edge_and_node_blocks_vs_oracle():
    """gather/concat, sum2, mean3 input assembly + mul + residual epilogue, ragged row counts."""
    from gnn_fluid_dynamics_b200 import ops, _lib
    g = torch.Generator().manual_seed(3)
    N, E, V = 333, 517, 190
    x, e = torch.randn(N, 128, generator=g), torch.randn(E, 128, generator=g)
    vs = torch.randn(V, 64, generator=g)
    mul = torch.randn(E, 128, generator=g)
    row, col = torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
    vf = [torch.randint(0, V, (N,), generator=g) for _ in range(3)]
    i32 = lambda t: t.to(torch.int32).to(dev())
    xd, ed, vsd = x.to(dev()), e.to(dev()), vs.to(dev())
    for prec in precisions():
        P_ = _lib.PRECISIONS[prec]
        # edge block, concat form
        p = _rand_mlp(384, 128, True, seed=5)
        ref = oracle.mlp3(torch.cat([e, x[row], x[col]], 1), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"])
        raw, summed = ops.mlp_forward([ops.Seg(ed), ops.Seg(xd, _lib.SEG_GATHER, (i32(row),)),
                                       ops.Seg(xd, _lib.SEG_GATHER, (i32(col),))], _to_weights(p, 0), E, P_,
                                      residual=ed, want_raw=True, want_sum=True)
        assert rel_l2(raw, ref) < TOL[prec] and rel_l2(summed, e + ref) < TOL[prec]
        # edge block, sum form with asym multiply
        p = _rand_mlp(256, 128, True, seed=6)
        ref = oracle.mlp3(torch.cat([e, x[row] + x[col]], 1), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"]) * mul
        raw, summed = ops.mlp_forward([ops.Seg(ed), ops.Seg(xd, _lib.SEG_SUM2, (i32(row), i32(col)))],
                                      _to_weights(p, 0), E, P_, mul=mul.to(dev()), residual=ed,
                                      want_raw=True, want_sum=True)
        assert rel_l2(raw, ref) < TOL[prec] and rel_l2(summed, e + ref) < TOL[prec]
        # node block, mean3 form
        p = _rand_mlp(192, 128, True, seed=7)
        agg = (vs[vf[0]] + vs[vf[1]] + vs[vf[2]]) / 3.0
        ref = oracle.mlp3(torch.cat([x, agg], 1), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"])
        raw, summed = ops.mlp_forward([ops.Seg(xd), ops.Seg(vsd, _lib.SEG_MEAN3, tuple(i32(t) for t in vf))],
                                      _to_weights(p, 0), N, P_, residual=xd, want_raw=True, want_sum=True)
        assert rel_l2(raw, ref) < TOL[prec] and rel_l2(summed, x + ref) < TOL[prec]


def _split_shadow(t, dtype):
    """hi | lo shadow of an fp32 matrix exactly as the kernels build it (round-to-nearest-even, lo = x - hi)."""
    hi = t.to(dtype)
    lo = (t - hi.float()).to(dtype)
    return torch.cat([hi, lo], 1).contiguous()


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
@pytest.mark.parametrize("E,N", [(1, 1), (33, 5), (1000, 700), (70001, 46000)])
def test_inference_fast_path_tma_gather_and_tma_store(prec, E, N):
    """The inference fast path of the fused block against the oracle AND against the register-staged path:
    gathered k-blocks staged by TMA gather4 from the 16-bit split shadow, final epilogue through TMA tensor stores
    (plain store, reduce-add into the in-place residual, load-add), the split shadow written by the epilogue.
    Ragged row counts (last tile clipped by the tensor map)."""
    from gnn_fluid_dynamics_b200 import ops, _lib
    from gnn_fluid_dynamics_b200.precisions import available
    if prec not in available():
        pytest.skip(prec)
    P_ = _lib.PRECISIONS[prec]
    sdt = ops.split_dtype(P_)
    g = torch.Generator().manual_seed(11)
    x, e = torch.randn(N, 128, generator=g), torch.randn(E, 128, generator=g)
    row, col = torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
    i32 = lambda t: t.to(torch.int32).to(dev())
    xd, ed = x.to(dev()), e.to(dev())
    p = _rand_mlp(384, 128, True, seed=5)
    w = _to_weights(p, 0)
    ref = oracle.mlp3(torch.cat([e, x[row], x[col]], 1), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"])
    # register-staged gathers, separate output buffers (the pre-existing path)
    segs = [ops.Seg(ed), ops.Seg(xd, _lib.SEG_GATHER, (i32(row),)), ops.Seg(xd, _lib.SEG_GATHER, (i32(col),))]
    raw0, sum0 = ops.mlp_forward(segs, w, E, P_, residual=ed, want_raw=True, want_sum=True)
    assert rel_l2(raw0, ref) < TOL[prec]
    # TMA-gathered from the shadow (the fp32 x is not passed at all), in-place residual, raw + sum
    xs = _split_shadow(xd, sdt)
    fsegs = lambda e_buf: [ops.Seg(e_buf), ops.Seg(xs.view(torch.float32), _lib.SEG_GATHER, (i32(row),), split=xs),
                           ops.Seg(xs.view(torch.float32), _lib.SEG_GATHER, (i32(col),), split=xs)]
    e1 = ed.clone()
    raw1, sum1 = ops.mlp_forward(fsegs(e1), w, E, P_, residual=e1, want_raw=True, want_sum=True, out_sum=e1)
    assert sum1.data_ptr() == e1.data_ptr()
    assert rel_l2(raw1, ref) < TOL[prec] and rel_l2(e1, e + ref) < TOL[prec]
    assert rel_l2(raw1, raw0) < 2e-6 and rel_l2(e1, sum0) < 2e-6      # same operands; only the epilogue's rounding differs
    # sum only, in place (reduce-add)
    e2 = ed.clone()
    ops.mlp_forward(fsegs(e2), w, E, P_, residual=e2, want_raw=False, want_sum=True, out_sum=e2)
    assert torch.equal(e2, e1)
    # sum only, separate buffer (load-add + plain store), with the split shadow of the SUM
    e3 = ed.clone()
    sh = torch.empty(E, 256, dtype=sdt, device=dev())
    _, sum3 = ops.mlp_forward(fsegs(e3), w, E, P_, residual=e3, want_raw=False, want_sum=True, out_split=sh, split_of_sum=True)
    assert torch.equal(e3, ed) and rel_l2(sum3, sum0) < 2e-6
    assert torch.equal(sh, _split_shadow(sum3, sdt))
    # shadow of the RAW output next to the in-place sum (the fvgn node block's hand-over), no fp32 raw written
    e4 = ed.clone()
    sh2 = torch.empty(E, 256, dtype=sdt, device=dev())
    ops.mlp_forward(fsegs(e4), w, E, P_, residual=e4, want_raw=False, want_sum=True, out_sum=e4, out_split=sh2)
    assert torch.equal(e4, e1) and torch.equal(sh2, _split_shadow(raw1, sdt))
    # determinism
    e5 = ed.clone()
    ops.mlp_forward(fsegs(e5), w, E, P_, residual=e5, want_raw=False, want_sum=True, out_sum=e5)
    assert torch.equal(e5, e2)


# --------------------------------------------------------------------- processors vs golden + oracle
def _run_processor(name, model, graphs_dev):
    from gnn_fluid_dynamics_b200.topology import get_topology
    c, f, v = graphs_dev
    grab = {}
    hook = lambda i, x, e: grab.__setitem__(i, (x, e)) if i in (0, 14) else None
    if name in ("ConservativeA", "ConservativeB", "ConservativeD", "ConservativeH", "ConservativeJ", "ConservativeK"):
        topo = get_topology(graphs_dev, need_cell_csr=True, two_hop=name in ("ConservativeH", "ConservativeJ", "ConservativeK")).validate()
        x, e, dec = model.encode_process_decode(c.x, f.x_symm, f.x_asym, topo, hook=hook)
        return {"x": x, "e": e, "dec": dec, "b1": grab[0]}
    topo = get_topology(graphs_dev, need_cell_csr=name.startswith("Conservative")).validate()
    if name == "ConservativeI":
        keep = ~((f.type == 2) | (f.type == 1)).reshape(-1)
        e_keep = keep.float().unsqueeze(1).expand(-1, 128).contiguous()
        x, e, dec = model.encode_process_decode(c.x, f.x, topo, hook=hook, e_keep=e_keep)
        return {"x": x, "e": e, "dec": dec, "b1": grab[0]}
    if name.startswith("VertPot"):
        x, e, vx, dec, dec_v = model.encode_process_decode(c.x, f.x, topo, hook=hook)
        return {"x": x, "e": e, "vx": vx, "dec": dec, "dec_vertex": dec_v, "b1": grab[0]}
    x, e, dec = model.encode_process_decode(c.x, f.x, topo, hook=hook)
    return {"x": x, "e": e, "dec": dec, "b1": grab[0]}


@pytest.mark.parametrize("name", ALL_MODELS)
def test_processor_matches_reference_golden(name):
    gold = load_golden(f"fwd_{name}.npz")
    model = build_model(name).eval()
    _, graphs = golden_graphs(name)
    graphs = [g.clone() for g in graphs]
    if name != "StreamFuncC":      # StreamFuncC.forward does not normalise (StreamFunc.py:173-176)
        graphs = model.normalizer.input(graphs)
    model.to(dev())
    gd = [g.to(dev()) for g in graphs]
    for prec in precisions():
        model.set_precision(prec)
        with torch.no_grad():
            out = _run_processor(name, model, gd)
        t = TOL[prec]
        assert rel_l2(out["b1"][0], torch.from_numpy(gold["x1"])) < t
        assert rel_l2(out["b1"][1], torch.from_numpy(gold["e1"])) < t
        assert rel_l2(out["x"], torch.from_numpy(gold["x15"])) < t, (prec, rel_l2(out["x"], torch.from_numpy(gold["x15"])))
        assert rel_l2(out["e"], torch.from_numpy(gold["e15"])) < t, (prec, rel_l2(out["e"], torch.from_numpy(gold["e15"])))
        assert rel_l2(out["dec"], torch.from_numpy(gold["dec"])) < 2 * t
        if name in ("ConservativeD", "ConservativeH", "ConservativeJ", "ConservativeK"):      # second (antisymmetric) edge stream
            assert rel_l2(model._last_e_asym, torch.from_numpy(gold["ea15"])) < t
        if name.startswith("VertPot"):
            assert rel_l2(out["vx"], torch.from_numpy(gold["vx15"])) < t
            assert rel_l2(out["dec_vertex"], torch.from_numpy(gold["dec_vertex"])) < 2 * t


@pytest.mark.parametrize("name", ALL_MODELS)
@pytest.mark.parametrize("mode", ["train", "rollout"])
def test_full_forward_matches_reference_golden(name, mode):
    gold = load_golden(f"fwd_{name}.npz")
    model = build_model(name).to(dev()).eval()
    _, graphs = golden_graphs(name)
    for prec in precisions():
        model.set_precision(prec)
        with torch.no_grad():
            out = model([g.clone().to(dev()) for g in graphs], mode=mode)
        for k, v in out.items():
            err = rel_l2(v, torch.from_numpy(gold[f"out_{mode}_{k}"]))
            assert err < 3 * TOL[prec], (name, k, prec, err)


@pytest.mark.parametrize("name", LOSS_MODELS)
def test_loss_matches_reference_golden(name):
    """model.loss(model(batch, 'train'), batch) of the glue-only variants against the value the reference computes."""
    gold = load_golden(f"fwd_{name}.npz")
    model = build_model(name).to(dev()).eval()
    _, graphs = golden_graphs(name)
    batch = [g.clone().to(dev()) for g in graphs]
    with torch.no_grad():
        losses = model.loss(model(batch, mode="train"), batch)
    assert set(f"loss_{k}" for k in losses) == set(k for k in gold if k.startswith("loss_"))
    for k, v in losses.items():
        ref = float(gold[f"loss_{k}"][0])
        assert abs(float(v) - ref) <= 2e-3 * max(abs(ref), 1e-6), (name, k, float(v), ref)


@pytest.mark.parametrize("name,n_cells", [("FvgnA", 20000), ("MgnA", 2048), ("FluxA", 20000), ("ConservativeA", 20000),
                                          ("ConservativeD", 20000)])
def test_processor_vs_oracle_at_baseline_sizes(name, n_cells):
    """BASELINE.json configs[0] (2k-cell MGN), the per-mesh size of configs[1] (20k-cell FVGN) and the config-3
    families (Flux / Conservative face-flux message passing) at 20k cells: every processor output within 1e-3."""
    model = build_model(name).eval()
    _, graphs = golden_graphs(name, n_cells=n_cells, mesh_seed=7, feat_seed=8)
    graphs = model.normalizer.input([g.clone() for g in graphs])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    c, f, v = graphs
    topo = {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face, "n_vertices": v.num_nodes}
    dual = name in ("ConservativeA", "ConservativeD")
    with torch.no_grad():
        ref = oracle.processor_fwd(oracle.family_of(name), sd, c.x, f.x_symm if dual else f.x, topo, 15,
                                   f_x_asym=f.x_asym if dual else None)
    model.to(dev())
    gd = [g.to(dev()) for g in graphs]
    for prec in precisions():
        model.set_precision(prec)
        with torch.no_grad():
            out = _run_processor(name, model, gd)
        assert rel_l2(out["x"], ref["x"]) < TOL[prec], (prec, rel_l2(out["x"], ref["x"]))
        assert rel_l2(out["e"], ref["e"]) < TOL[prec], (prec, rel_l2(out["e"], ref["e"]))
        assert rel_l2(out["dec"], ref["dec"]) < 2 * TOL[prec], (prec, rel_l2(out["dec"], ref["dec"]))
        if name == "ConservativeD":
            assert rel_l2(model._last_e_asym, ref["ea"]) < TOL[prec]


def test_batched_meshes_equal_individual_meshes():
    """PyG-style concatenated batch == per-mesh results (independent units)."""
    from gnn_fluid_dynamics_b200.graph import collate_triplet
    from gnn_fluid_dynamics_b200.topology import get_topology
    model = build_model("FvgnA").to(dev()).eval()
    samples = [golden_graphs("FvgnA", n_cells=n, mesh_seed=s)[1] for n, s in [(300, 1), (500, 2), (200, 3)]]
    with torch.no_grad():
        singles = []
        for s in samples:
            gd = [g.clone().to(dev()) for g in s]
            singles.append(model.encode_process_decode(gd[0].x, gd[1].x, get_topology(gd).validate()))
        batch = [g.to(dev()) for g in collate_triplet([[g.clone() for g in s] for s in samples])]
        xb, eb, db = model.encode_process_decode(batch[0].x, batch[1].x, get_topology(batch).validate())
    assert torch.equal(xb, torch.cat([s[0] for s in singles]))     # row-independent kernels: bitwise
    assert torch.equal(eb, torch.cat([s[1] for s in singles]))
    assert torch.equal(db, torch.cat([s[2] for s in singles]))


def test_determinism_bitwise_repeatable():
    from gnn_fluid_dynamics_b200.topology import get_topology
    model = build_model("MgnA").to(dev()).eval()
    _, graphs = golden_graphs("MgnA", n_cells=5000)
    gd = [g.to(dev()) for g in graphs]
    topo = get_topology(gd).validate()
    with torch.no_grad():
        a = model.encode_process_decode(gd[0].x, gd[1].x, topo)
        b = model.encode_process_decode(gd[0].x, gd[1].x, topo)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


@pytest.mark.parametrize("prec", ["f32", "bf16x3"])
def test_signed_sum3_segment_equals_cell_signed_sum(prec):
    """SEG_SUM3S (the Conservative models' signed edge->cell aggregation as the node MLP's own input assembly,
    Conservative.py:243-254) against the segment-sum kernel + a DIRECT segment: interior cells bit-identical sums, the MLP
    outputs within the precision's tolerance; and ConservativeA's forward with the fused path against the unfused one."""
    from gnn_fluid_dynamics_b200 import ops, _lib, processor as P
    from gnn_fluid_dynamics_b200.ops import Seg
    from gnn_fluid_dynamics_b200.topology import get_topology
    dev = torch.device("cuda:0")
    _, graphs = golden_graphs("ConservativeA", n_cells=3000, mesh_seed=7, feat_seed=8)
    gd = [g.to(dev) for g in graphs]
    topo = get_topology(gd, need_cell_csr=True, two_hop=False).validate()
    ell = topo.build_signed_cell_ell(gd[1].face)
    N, E = gd[0].x.shape[0], gd[0].edge_index.shape[1]
    gen = torch.Generator().manual_seed(3)
    e_raw = torch.randn(E, 128, generator=gen).to(dev)
    x = torch.randn(N, 128, generator=gen).to(dev)
    agg = P.cell_signed_sum(e_raw, topo)
    # the table itself: decode and sum on the host side of the device
    ref = torch.zeros_like(agg)
    for i in ell:
        i = i.long()
        zero = i == _lib.SUM3S_ZERO
        rows = torch.where(i >= 0, i, -i - 1).clamp(min=0)
        sign = torch.where(zero, 0.0, torch.where(i >= 0, 1.0, -1.0)).to(torch.float32)
        ref = ref + sign[:, None] * e_raw[torch.where(zero, torch.zeros_like(rows), rows)]
    assert rel_l2(ref, agg) < 1e-6
    w = _to_weights(_rand_mlp(256, 128, True, seed=4), 0)
    p = _lib.PRECISIONS[prec]
    fused, _ = ops.mlp_forward([Seg(x), Seg(e_raw, _lib.SEG_SUM3S, ell)], w, N, p)
    plain, _ = ops.mlp_forward([Seg(x), Seg(agg)], w, N, p)
    assert rel_l2(fused, plain) < (1e-6 if prec == "f32" else 2e-5), rel_l2(fused, plain)
    model = build_model("ConservativeA", precision=prec if prec != "f32" else None).to(dev).eval()
    with torch.no_grad():
        out_p = model([g.clone() for g in gd], mode="rollout")
        P.FUSE_SIGNED_SUM = True
        try:
            out_f = model([g.clone() for g in gd], mode="rollout")
        finally:
            P.FUSE_SIGNED_SUM = False
    for k in out_f:
        assert rel_l2(out_f[k], out_p[k]) < 1e-4, (k, rel_l2(out_f[k], out_p[k]))
