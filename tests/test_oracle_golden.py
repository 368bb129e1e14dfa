"""The oracle (oracle/) against the fixtures produced by the reference itself (tests/golden/*.npz):
this is what pins the oracle (SURVEY.md section 8c - the reference ships no golden vectors)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import model as omodel
from gnn_fluid_dynamics_b200.mesh import connectivity
from fixtures import default_stats, rel_l2
from helpers import ALL_MODELS, GOLDEN, LOSS_W, build_model, golden_graphs, load_golden

MODELS = ALL_MODELS
TOL = 2e-5   # fp32 CPU restatement vs fp32 CPU reference: summation-order noise only


@pytest.mark.parametrize("i", [0, 1, 2])
def test_connectivity_matches_reference(i):
    g = load_golden(f"connectivity_{i}.npz")
    fi, cei, vei = connectivity(g["cells"], g["vertex_pos"])
    assert np.array_equal(fi, g["face_index"])
    assert np.array_equal(cei, g["cell_edge_index"])
    assert np.array_equal(vei, g["vertex_edge_index"])


@pytest.mark.parametrize("name", MODELS)
def test_state_dict_keys_match_reference(name):
    ref = json.load(open(os.path.join(GOLDEN, f"keys_{name}.json")))
    mine = [[k, list(v.shape)] for k, v in build_model(name).state_dict().items()]
    assert mine == ref


@pytest.mark.parametrize("name", MODELS)
def test_update_features_matches_reference(name):
    """The rollout feature update (plain tensor glue, runs on CPU) against the reference's own update_features."""
    gold = load_golden(f"upd_{name}.npz")
    model = build_model(name)
    _, graphs = golden_graphs(name)
    g = [x.clone() for x in graphs]
    g[1].type = g[1].type.reshape(-1)
    out = {"cell_velocity": torch.randn(g[0].x.shape[0], 2, generator=torch.Generator().manual_seed(31))}
    c, f, v = model.update_features(out, g)
    fx = f.x_asym if hasattr(f, "x_asym") else f.x
    assert torch.equal(c.x, torch.from_numpy(gold["cx"]))
    assert torch.equal(fx[:, 0:2], torch.from_numpy(gold["fx2"]))


@pytest.mark.parametrize("name", MODELS)
def test_class_interface_matches_reference(name):
    """Flags, feature sizes and normalisation registry the reference's training / dataset code reads off the class."""
    ref = json.load(open(os.path.join(GOLDEN, f"meta_{name}.json")))
    model = build_model(name)
    cls = type(model)
    for a, v in ref["flags"].items():
        assert bool(getattr(model, a, False)) == v, a
    assert [a for a in ("cell_mls_weights", "face_mls_weights") if hasattr(model, a)] == ref["mls_attrs"]
    assert [list(x) for x in cls.get_feature_sizes(None)] == ref["feature_sizes"]
    reg, ins, outs = cls.get_normalisation_map()
    assert {k: v[1] for k, v in reg.items()} == ref["registry_kinds"]
    assert sorted(v[1] for v in ins.values()) == ref["input_stat_keys"]
    assert sorted(v[1] for v in outs.values()) == ref["output_stat_keys"]


@pytest.mark.parametrize("name", ["VertPotD", "VertPotF"])
def test_unrunnable_reference_classes_construct_and_explain(name):
    """The two reference classes whose forward raises in the reference itself (missing fvm function): same constructor
    and state_dict layout (checkpoints load), forward raises with the reason instead of an AttributeError."""
    from gnn_fluid_dynamics_b200.models import UNRUNNABLE_CLASSES
    from fixtures import stats_for
    from helpers import make_config, mse
    model = UNRUNNABLE_CLASSES[name](make_config(), mse, None, stats_for(name))
    ref = json.load(open(os.path.join(GOLDEN, f"keys_{name}.json")))
    assert [[k, list(v.shape)] for k, v in model.state_dict().items()] == ref
    with pytest.raises(NotImplementedError, match="convert_cell_flux_to_face_flux_alt"):
        model(None)


TRANSFORM_MODELS = MODELS


@pytest.mark.parametrize("name", TRANSFORM_MODELS)
def test_transform_features_matches_reference(name):
    """cls.transform_features on raw series (train mode: seeded noise + edge flip; valid mode) vs the reference."""
    from types import SimpleNamespace as NS
    from gnn_fluid_dynamics_b200.mesh import make_mesh
    from gnn_fluid_dynamics_b200.models import MODEL_CLASSES
    from fixtures import raw_graphs
    from helpers import GOLDEN_SETUP
    gold = load_golden(f"tf_{name}.npz")
    types = ["NORMAL", "WALL_BOUNDARY", "INFLOW", "OUTFLOW", "SLIP"]
    for tag, noise, mode in (("train", True, "train"), ("valid", False, "valid")):
        ds = NS(class_types=types, noise=noise, mode=mode, config=NS(training=NS(noise_std=0.02)))
        graphs = raw_graphs(make_mesh(160, GOLDEN_SETUP[name][0], seed=4))
        torch.manual_seed(77)
        c, f, v = MODEL_CLASSES[name].transform_features(ds, graphs)
        seen = 0
        for gname, g in (("c", c), ("f", f)):
            for k in ("x", "y", "x_symm", "x_asym", "edge_index", "normal", "boundary_mask", "flux"):
                key = f"{tag}_{gname}_{k}"
                has = hasattr(g, k) and torch.is_tensor(getattr(g, k))
                assert has == (key in gold), key
                if not has:
                    continue
                mine, ref = getattr(g, k), torch.from_numpy(gold[key])
                assert mine.shape == ref.shape and mine.dtype == ref.dtype, key
                if mine.is_floating_point():
                    assert torch.allclose(mine, ref, rtol=1e-6, atol=1e-6), key
                else:
                    assert torch.equal(mine, ref), key
                seen += 1
        assert seen >= 6
        assert int(hasattr(c, "velocity")) == int(gold[f"{tag}_c_has_velocity"][0])


def _processor_inputs(name, graphs, model):
    if name != "StreamFuncC":      # StreamFuncC.forward does not normalise (StreamFunc.py:173-176)
        graphs = model.normalizer.input(graphs)
    c, f, v = graphs
    topo = {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face,
            "n_vertices": v.num_nodes}
    if name in ("ConservativeA", "ConservativeB", "ConservativeD", "ConservativeH", "ConservativeJ", "ConservativeK"):
        return c.x, f.x_symm, f.x_asym, topo
    return c.x, f.x, None, topo


@pytest.mark.parametrize("name", MODELS)
def test_oracle_processor_matches_reference(name):
    gold = load_golden(f"fwd_{name}.npz")
    model = build_model(name)
    sd = model.state_dict()
    _, graphs = golden_graphs(name)
    c_x, f_x, f_xa, topo = _processor_inputs(name, [g.clone() for g in graphs], model)
    fam = oracle.family_of(name)
    with torch.no_grad():
        bc = ((graphs[1].type == 2) | (graphs[1].type == 1)).reshape(-1) if name == "ConservativeI" else None
        out = oracle.processor_fwd(fam, sd, c_x, f_x, topo, 15, f_x_asym=f_xa, keep_blocks=True, bc_mask=bc)
    assert rel_l2(out["x0"], torch.from_numpy(gold["x0"])) < TOL
    assert rel_l2(out["e0"], torch.from_numpy(gold["e0"])) < TOL
    assert rel_l2(out["blocks"][0][0], torch.from_numpy(gold["x1"])) < TOL
    assert rel_l2(out["blocks"][0][1], torch.from_numpy(gold["e1"])) < TOL
    assert rel_l2(out["x"], torch.from_numpy(gold["x15"])) < TOL
    assert rel_l2(out["e"], torch.from_numpy(gold["e15"])) < TOL
    if name.startswith("VertPot"):
        assert rel_l2(out["vx"], torch.from_numpy(gold["vx15"])) < TOL
        assert rel_l2(out["dec"][0], torch.from_numpy(gold["dec"])) < TOL
        assert rel_l2(out["dec"][1], torch.from_numpy(gold["dec_vertex"])) < TOL
    else:      # (FvgnC's decoder output is viewed [E, k, 5], Fvgn.py:780-786)
        assert rel_l2(out["dec"].reshape(gold["dec"].shape), torch.from_numpy(gold["dec"])) < TOL


@pytest.mark.parametrize("name", ["FvgnA", "MgnA"])
@pytest.mark.parametrize("mode", ["train", "rollout"])
def test_oracle_full_forward_matches_reference(name, mode):
    gold = load_golden(f"fwd_{name}.npz")
    sd = build_model(name).state_dict()
    _, graphs = golden_graphs(name)
    with torch.no_grad():
        out, _ = omodel.model_forward(name, sd, default_stats(), [g.clone() for g in graphs], 15, mode=mode)
    for k, v in out.items():
        assert rel_l2(v, torch.from_numpy(gold[f"out_{mode}_{k}"])) < TOL, k


def test_oracle_train_step_matches_reference():
    gold = load_golden("train_FvgnA.npz")
    model = build_model("FvgnA")
    params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    _, graphs = golden_graphs("FvgnA", flip=True)
    out, _ = omodel.model_forward("FvgnA", params, default_stats(), [g.clone() for g in graphs], 15,
                                  mode="train", training=True)
    graphs_n = omodel.normalise_inputs("FvgnA", default_stats(), [g.clone() for g in graphs])
    losses = omodel.fvgn_loss(params, out, graphs_n, LOSS_W, training=True)
    for k, v in losses.items():
        assert abs(float(v) - float(gold[f"loss_{k}"][0])) < 1e-4 * max(1.0, abs(float(gold[f"loss_{k}"][0]))), k
    losses["total_log_loss"].backward()
    names = [str(n) for n in gold["grad_names"]]
    for n, ref_norm in zip(names, gold["grad_norms"]):
        g = params[n].grad
        assert g is not None, n
        assert abs(float(g.double().norm()) - ref_norm) <= 2e-4 * max(ref_norm, 1e-6) + 1e-9, n
    for k in gold:
        if k.startswith("grad_processer") or k.startswith("grad_decoder") or k.startswith("grad_encoder"):
            assert rel_l2(params[k[5:]].grad, torch.from_numpy(gold[k])) < 1e-4, k


def test_scatter_add_matches_loop_and_csr_is_stable_sort():
    rng = np.random.RandomState(0)
    idx = rng.randint(0, 37, size=400)
    src = rng.randn(400, 8).astype(np.float32)
    a = oracle.scatter_add(torch.from_numpy(src), torch.from_numpy(idx), 40).numpy()
    b = oracle.scatter_add_loop(src, idx, 40)
    assert np.array_equal(a, b)       # same summation order -> bit-identical
    off, perm = oracle.csr_build(idx, 40)
    ref_perm = torch.sort(torch.from_numpy(idx), stable=True).indices.numpy()
    assert np.array_equal(perm, ref_perm)
    assert np.array_equal(np.diff(off), np.bincount(idx, minlength=40))


@pytest.mark.parametrize("name", ["FvgnA", "MgnA"])
def test_oracle_rollout_matches_reference_rollout(name):
    """oracle.model.rollout_step against the reference's own 100-step loop (rollout_{name}.npz, generated by
    tests/golden/make_golden.py --rollout-only from src/rollout.py:313-369 semantics)."""
    from oracle import model as omodel
    from fixtures import default_stats
    gold = load_golden(f"rollout_{name}.npz")
    model = build_model(name).eval()
    _, graphs = golden_graphs(name, n_cells=400, mesh_seed=31, feat_seed=32)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    graphs = [g.clone() for g in graphs]
    with torch.no_grad():
        for step in range(1, 11):
            vel = omodel.rollout_step(name, sd, default_stats(), graphs, 15)
            if step in (1, 10):
                assert rel_l2(vel, torch.from_numpy(gold[f"vel_{step}"])) < 1e-4, (name, step)


@pytest.mark.parametrize("name", ["FvgnA", "MgnA", "FluxA", "ConservativeA", "VertPotA"])
def test_state_dict_from_keys_equals_constructed_model(name):
    """fixtures.state_dict_from_keys (what bench.py's CPU reference arm uses, so that it never loads the CUDA library)
    reproduces the constructed model's deterministic state_dict bit for bit."""
    from fixtures import state_dict_from_keys
    a = state_dict_from_keys(name)
    b = build_model(name).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k].float(), b[k].float()), k
