"""Generate the committed golden fixtures by running the REFERENCE's own code in this container.

Run here only:  python tests/golden/make_golden.py
(needs /root/reference, which does not exist on the GPU box; the .npz files it writes travel.)

What is pinned
  connectivity_{k}.npz : reference utils/geometry.compute_connectivity on three synthetic meshes
  fwd_{Model}.npz      : reference model forward (every class in MODELS below: 36 of the reference's 38; VertPotD / F
                         cannot run in the reference),
                         hidden 128, 15 blocks, deterministic parameters
                         (tests/fixtures.py fill_state_dict_deterministic, seed 1),
                         mesh make_mesh(160, kind, seed=3), features mesh_graphs(seed=5):
                         encoder outputs, processor outputs after block 1 and block 15, decoder
                         outputs, the forward() dict in 'train' and 'rollout' modes; for LOSS_MODELS also
                         model.loss(forward(batch, 'train'), batch)
  train_FvgnA.npz      : reference FvgnA train-mode forward + model.loss + backward: loss values and
                         gradients (norm of every parameter's grad + a few full tensors)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))      # tests/ (fixtures.py)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))   # repo root
import refstub  # noqa: E402

refstub.install()

from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs  # noqa: E402
from fixtures import add_mls_fixture, default_stats, fill_state_dict_deterministic, raw_graphs, stats_for  # noqa: E402
from gnn_fluid_dynamics_b200.graph import Data  # noqa: E402

from utils.config import Config  # noqa: E402  (reference)
from utils.geometry import compute_connectivity  # noqa: E402
from utils.loss import MSE_per_element_torch  # noqa: E402
from datasets.OpenFoam import NodeType  # noqa: E402
import importlib  # noqa: E402

torch.set_num_threads(4)

MODELS = {
    "MgnA": ("models.Mgn", "cylinder", "fvgn"),
    "FvgnA": ("models.Fvgn", "cylinder", "fvgn"),
    "FluxA": ("models.Flux", "ellipse", "fvgn"),
    "ConservativeA": ("models.Conservative", "cylinder", "conservative"),
    "VertPotA": ("models.VertPot", "airfoil", "fvgn"),
    "ConservativeE": ("models.Conservative", "ellipse", "fvgn"),
    "ConservativeF": ("models.Conservative", "airfoil", "fvgn"),
    "ConservativeD": ("models.Conservative", "ellipse", "conservative"),
    "ConservativeG": ("models.Conservative", "cylinder", "fvgn"),
    "ConservativeI": ("models.Conservative", "airfoil", "fvgn"),
    "ConservativeH": ("models.Conservative", "cylinder", "conservative_h"),
    "FvgnF": ("models.Fvgn", "airfoil", "fvgn"),
    "ConservativeK": ("models.Conservative", "ellipse", "conservative_h"),
    "MgnB": ("models.Mgn", "ellipse", "fvgn"),
    "MgnC": ("models.Mgn", "airfoil", "fvgn"),
    "StreamFuncA": ("models.StreamFunc", "cylinder", "fvgn"),
    "StreamFuncB": ("models.StreamFunc", "ellipse", "fvgn"),
    "StreamFuncC": ("models.StreamFunc", "airfoil", "fvgn"),
    "StreamFuncD": ("models.StreamFunc", "cylinder", "fvgn"),
    "ConservativeB": ("models.Conservative", "airfoil", "conservative"),
    "ConservativeJ": ("models.Conservative", "ellipse", "conservative_h"),
    "VertPotB": ("models.VertPot", "cylinder", "fvgn"),
    "VertPotC": ("models.VertPot", "ellipse", "fvgn"),
    "VertPotE": ("models.VertPot", "airfoil", "fvgn"),
    "VertPotG": ("models.VertPot", "cylinder", "fvgn"),
    "FvgnC": ("models.Fvgn", "ellipse", "fvgn"),
    "FluxB": ("models.Flux", "cylinder", "fvgn"),
    "FluxC": ("models.Flux", "airfoil", "fvgn"),
    "FluxD": ("models.Flux", "ellipse", "fvgn"),
    "FvgnB": ("models.Fvgn", "cylinder", "fvgn"),
    "FvgnD": ("models.Fvgn", "ellipse", "fvgn"),
    "FvgnE": ("models.Fvgn", "airfoil", "fvgn"),
    "FvgnH": ("models.Fvgn", "cylinder", "fvgn"),
    "FvgnI": ("models.Fvgn", "ellipse", "fvgn"),
    "FvgnJ": ("models.Fvgn", "airfoil", "fvgn"),
    "FvgnK": ("models.Fvgn", "cylinder", "fvgn"),
}
FVGN_LIKE = ("FvgnA", "FvgnB", "FvgnD", "FvgnE", "FvgnH", "FvgnI", "FvgnJ", "FvgnK")
MGN_LIKE = ("MgnA", "MgnB", "MgnC", "StreamFuncA", "StreamFuncB", "StreamFuncC", "StreamFuncD")
LOSS_W = {"continuity": 0, "cell_velocity_change": 10, "cell_pressure": 1, "face_velocity": 1,
          "face_flux": 1, "face_pressure": 1, "cell_velocity": 10}
# models whose fixture also pins model.loss(forward(batch, 'train'), batch) (eval mode, no grad)
LOSS_MODELS = ("FvgnC", "VertPotC", "VertPotE", "VertPotG", "ConservativeB", "ConservativeJ", "FvgnB", "FvgnE", "FvgnH", "FvgnJ", "FvgnK", "MgnB", "MgnC", "StreamFuncA", "StreamFuncB", "StreamFuncC", "StreamFuncD", "FluxB", "FluxC", "FluxD")


class _Dataset:
    class_types = NodeType
    noise = False
    mode = "valid"


def ref_config():
    return Config.from_dict({
        "model": {"hidden_width": 128, "mp_num": 15, "bundle_size": 3, "cell_grad_weights_order": 1,
                  "face_grad_weights_order": 1},
        "training": {"dropout_rate": 0.0, "loss_weights": LOSS_W},
    })


def build_ref(name):
    module, kind, flavour = MODELS[name]
    cls = getattr(importlib.import_module(module), name)
    model = cls(ref_config(), MSE_per_element_torch, _Dataset(), stats_for(name))
    fill_state_dict_deterministic(model, seed=1)
    return model, kind, flavour


def graphs_for(name, kind, flavour, flip=False, n_cells=160, mesh_seed=3, feat_seed=5):
    mesh = make_mesh(n_cells, kind, seed=mesh_seed)
    g = mesh_graphs(mesh, seed=feat_seed, flavour=flavour, flip_edges=flip)
    c, f, v = g
    if name in MGN_LIKE + ("ConservativeB",):
        c.y = torch.cat([c.y, torch.randn(c.x.shape[0], 1, generator=torch.Generator().manual_seed(9))], 1)
        f.y = f.y[:, :2].contiguous()
    elif name in FVGN_LIKE + ("ConservativeA", "VertPotA", "ConservativeE", "ConservativeF", "ConservativeD", "ConservativeG", "ConservativeI", "ConservativeH", "FvgnF", "ConservativeK", "ConservativeJ"):
        f.y = f.y[:, :3].contiguous() if name != "VertPotA" else f.y
    if name == "FluxC":
        f.y = f.y[:, :2].contiguous()
    if name in ("VertPotC", "VertPotE"):
        f.y = f.y[:, :2].contiguous()
    if name in ("FvgnB", "VertPotB"):
        add_mls_fixture(f, seed=12)
    if name == "FvgnH":
        extra = torch.randn(f.x.shape[0], 2, generator=torch.Generator().manual_seed(13))
        f.x = torch.cat([f.x[:, :5], extra, f.x[:, 5:]], dim=1)
    if name == "FvgnC":
        c.y = torch.randn(c.x.shape[0], 3, 2, generator=torch.Generator().manual_seed(21))
        f.y = torch.randn(f.x.shape[0], 3, 3, generator=torch.Generator().manual_seed(22))
    if name == "FvgnK":
        c.Re = torch.tensor([150.0])
        f.type = f.type.reshape(-1)
    if name == "ConservativeI":
        # the reference indexes the [E, 128] latent with the face-type mask (Conservative.py:1264-1267), which only
        # works for a 1-D type tensor
        f.type = f.type.reshape(-1)
    if name.startswith("StreamFunc") or name in ("MgnB", "MgnC"):
        add_mls_fixture(c)
    c.batch = torch.zeros(c.x.shape[0], dtype=torch.long)
    f.batch = torch.zeros(f.pos.shape[0], dtype=torch.long)
    return mesh, g


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items() if torch.is_tensor(v)}


def gen_connectivity():
    for i, (n, kind, srt) in enumerate([(200, "cylinder", False), (700, "airfoil", True), (512, "none", False)]):
        m = make_mesh(n, kind, seed=i, sort_cell_vertices=srt)
        fi, cei, vei = compute_connectivity(m.cells, m.vertex_pos)
        np.savez_compressed(os.path.join(HERE, f"connectivity_{i}.npz"), cells=m.cells,
                            vertex_pos=m.vertex_pos, face_index=fi, cell_edge_index=cei,
                            vertex_edge_index=vei)
        print("connectivity", i, m.n_cells, m.n_faces, m.n_vertices)


def gen_forward(name):
    model, kind, flavour = build_ref(name)
    model.eval()
    mesh, graphs = graphs_for(name, kind, flavour)
    cap = {}

    def grab_enc(mod, inp, out):
        cap["x0"], cap["e0"] = out.x.clone(), out.edge_attr.clone()
        if hasattr(out, "edge_attr_asym"):
            cap["e0_asym"] = out.edge_attr_asym.clone()

    def grab_block(tag):
        def hook(mod, inp, out):
            cg = out[0] if isinstance(out, tuple) else out
            cap[f"x{tag}"], cap[f"e{tag}"] = cg.x.clone(), cg.edge_attr.clone()
            if hasattr(cg, "edge_attr_asym"):
                cap[f"ea{tag}"] = cg.edge_attr_asym.clone()
            if isinstance(out, tuple):
                cap[f"vx{tag}"] = out[1].x.clone()
        return hook

    def grab_dec(mod, inp, out):
        if isinstance(out, tuple):
            cap["dec"], cap["dec_vertex"] = out[0].clone(), out[1].clone()
        else:
            cap["dec"] = out.clone()

    model.encoder.register_forward_hook(grab_enc)
    if hasattr(model, "gn_block"):      # FvgnF: ONE shared block applied mp_num times
        calls = {"n": 0}

        def grab_shared(mod, inp, out):
            calls["n"] = calls["n"] % 15 + 1
            if calls["n"] in (1, 15):
                grab_block(calls["n"])(mod, inp, out)
        model.gn_block.register_forward_hook(grab_shared)
    else:
        model.processer_list[0].register_forward_hook(grab_block(1))
        model.processer_list[-1].register_forward_hook(grab_block(15))
    model.decoder.register_forward_hook(grab_dec)

    out = {}
    with torch.no_grad():
        for mode in ("train", "rollout"):
            res = model([g.clone() for g in graphs], mode=mode)
            for k, v in res.items():
                out[f"out_{mode}_{k}"] = v.clone()
        if name in LOSS_MODELS:
            batch = [g.clone() for g in graphs]
            for k, v in model.loss(model(batch, mode="train"), batch).items():
                out[f"loss_{k}"] = v.reshape(1).clone()
    out.update(cap)
    np.savez_compressed(os.path.join(HERE, f"fwd_{name}.npz"), **to_np(out))
    print(name, {k: tuple(v.shape) for k, v in out.items()})


def gen_update(name):
    """Reference update_features (the rollout feature update, e.g. Fvgn.py:133-148) on the fixture graphs.  The
    reference indexes a [E, 2] tensor with the face-type mask, which needs a 1-D type tensor."""
    model, kind, flavour = build_ref(name)
    mesh, graphs = graphs_for(name, kind, flavour)
    g = [x.clone() for x in graphs]
    g[1].type = g[1].type.reshape(-1)
    n = g[0].x.shape[0]
    out = {"cell_velocity": torch.randn(n, 2, generator=torch.Generator().manual_seed(31))}
    c, f, v = model.update_features(out, g)
    fx = f.x_asym if hasattr(f, "x_asym") else f.x
    np.savez_compressed(os.path.join(HERE, f"upd_{name}.npz"), cx=c.x.numpy(), fx2=fx[:, 0:2].numpy())


TRANSFORM_MODELS = list(MODELS)      # 14 classes define it, the others inherit (resolved through the MRO)


class _TrainDataset:
    class_types = NodeType
    noise = True
    mode = "train"
    config = Config.from_dict({"model": {"hidden_width": 128, "mp_num": 15}, "training": {"noise_std": 0.02}})


def gen_transform(name):
    """Reference ``cls.transform_features(dataset, raw graphs)`` (train mode: noise + random edge flip under
    torch.manual_seed(77), and valid mode) on a 160-cell mesh."""
    module, kind, _ = MODELS[name]
    cls = getattr(importlib.import_module(module), name)
    out = {}
    for tag, ds in (("train", _TrainDataset()), ("valid", _Dataset())):
        graphs = raw_graphs(make_mesh(160, kind, seed=4))
        torch.manual_seed(77)
        c, f, v = cls.transform_features(ds, graphs)
        for gname, g in (("c", c), ("f", f)):
            for k in ("x", "y", "x_symm", "x_asym", "edge_index", "normal", "boundary_mask", "flux"):
                if hasattr(g, k) and torch.is_tensor(getattr(g, k)):
                    out[f"{tag}_{gname}_{k}"] = getattr(g, k).clone()
        out[f"{tag}_c_has_velocity"] = torch.tensor([int(hasattr(c, "velocity"))])
    np.savez_compressed(os.path.join(HERE, f"tf_{name}.npz"), **to_np(out))


def gen_keys(name):
    import json
    model, _, _ = build_ref(name)
    keys = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    json.dump(keys, open(os.path.join(HERE, f"keys_{name}.json"), "w"))
    # class-level interface the callers read (train.py:350-355, DataSet.py:318, 344-352)
    reg, ins, outs = type(model).get_normalisation_map()
    meta = {"flags": {a: bool(getattr(model, a, False)) for a in
                      ("pushforward_use", "cell_grad_weights_use", "face_grad_weights_use")},
            "mls_attrs": [a for a in ("cell_mls_weights", "face_mls_weights") if hasattr(model, a)],
            "feature_sizes": [list(x) for x in type(model).get_feature_sizes(_Dataset())],
            "registry_kinds": {k: v[1] for k, v in reg.items()},
            "input_stat_keys": sorted(v[1] for v in ins.values()),
            "output_stat_keys": sorted(v[1] for v in outs.values())}
    json.dump(meta, open(os.path.join(HERE, f"meta_{name}.json"), "w"), sort_keys=True)


def gen_train():
    name = "FvgnA"
    model, kind, flavour = build_ref(name)
    model.train()
    mesh, graphs = graphs_for(name, kind, flavour, flip=True)
    batch = [g.clone() for g in graphs]
    output = model(batch, mode="train")
    losses = model.loss(output, batch)
    losses["total_log_loss"].backward()
    out = {f"loss_{k}": v.detach().reshape(1) for k, v in losses.items()}
    keep_full = ["processer_list.0.cell_block.cell_mlp.1.weight", "processer_list.14.face_block.face_mlp.1.bias",
                 "decoder.face_mlp.4.weight", "encoder.cell_mlp.0.0.weight",
                 "processer_list.7.face_block.face_mlp.0.0.bias"]
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(float(p.grad.double().norm()))
        if k in keep_full:
            out["grad_" + k] = p.grad.clone()
    out["grad_norms"] = torch.tensor(norms, dtype=torch.float64)
    for k, v in output.items():
        out["out_" + k] = v.detach()
    d = to_np(out)
    d["grad_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "train_FvgnA.npz"), **d)
    print("train", {k: float(v) for k, v in losses.items()}, len(names), "param grads")


TRAIN_STEP_MODELS = ("VertPotA", "StreamFuncA", "FluxA", "ConservativeA", "MgnA")


def gen_train_step(name):
    """Reference training step (train.py:251-256: model.train(); forward(batch, 'train'); model.loss; backward) on
    the 160-cell fixture with flipped edges: loss values, the norm of every parameter gradient and up to six full
    gradient tensors.  VertPotA / StreamFuncA are BASELINE.json config 5's families."""
    model, kind, flavour = build_ref(name)
    model.train()
    _, graphs = graphs_for(name, kind, flavour, flip=True)
    batch = [g.clone() for g in graphs]
    output = model(batch, mode="train")
    losses = model.loss(output, batch)
    losses["total_log_loss"].backward()
    out = {f"loss_{k}": v.detach().reshape(1) for k, v in losses.items()}
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is None or float(p.grad.abs().max()) == 0.0:      # VertPot's unused duplicate blocks get no gradient
            continue
        names.append(k)
        norms.append(float(p.grad.double().norm()))
    keep = names[:: max(1, len(names) // 6)][:6]
    grads = dict(model.named_parameters())
    for k in keep:
        out["grad_" + k] = grads[k].grad.clone()
    out["grad_norms"] = torch.tensor(norms, dtype=torch.float64)
    d = to_np(out)
    d["grad_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, f"train_{name}.npz"), **d)
    print("train", name, {k: float(v.detach()) for k, v in losses.items()}, len(names), "param grads", keep)


ROLLOUT_MODELS = ("FvgnA", "MgnA", "FluxA", "ConservativeA", "ConservativeD", "MgnB", "StreamFuncA")
ROLLOUT_KEEP = (1, 10, 50, 100)


def gen_rollout(name):
    """The reference's autoregressive loop (src/rollout.py:313-369) run on CPU for 100 steps on a 400-cell mesh:
    output = model(clones, mode='rollout'); cell_velocity = output['cell_velocity'] if the model returns it, else
    x[:, :2] + cell_velocity_change; input_graphs = model.update_features(solutions, input_graphs).
    Same mesh / features as tests/test_gpu_rollout.py (n_cells=400, mesh_seed=31, feat_seed=32)."""
    model, kind, flavour = build_ref(name)
    model.eval()
    _, graphs = graphs_for(name, kind, flavour, n_cells=400, mesh_seed=31, feat_seed=32)
    graphs = [g.clone() for g in graphs]
    type2d = graphs[1].type.reshape(-1, 1).clone()
    out = {}
    with torch.no_grad():
        for step in range(1, 101):
            graphs[1].type = type2d
            sol = dict(model([g.clone() for g in graphs], mode="rollout"))
            if "cell_velocity" not in sol:
                sol["cell_velocity"] = graphs[0].x[:, 0:2] + sol["cell_velocity_change"]
            graphs[1].type = type2d.reshape(-1)          # the reference's boolean-mask assignment needs a 1-D type
            graphs = list(model.update_features(sol, graphs))
            if step in ROLLOUT_KEEP:
                out[f"vel_{step}"] = sol["cell_velocity"].clone()
    assert all(torch.isfinite(v).all() for v in out.values()), name
    np.savez_compressed(os.path.join(HERE, f"rollout_{name}.npz"), **to_np(out))
    print("rollout", name, {k: float(v.norm()) for k, v in out.items()})


if __name__ == "__main__":
    if sys.argv[1:2] == ["--train-only"]:
        for n in TRAIN_STEP_MODELS:
            gen_train_step(n)
        sys.exit(0)
    if sys.argv[1:2] == ["--rollout-only"]:
        for n in ROLLOUT_MODELS:
            gen_rollout(n)
        sys.exit(0)
    if sys.argv[1:2] == ["--update-only"]:
        for n in MODELS:
            gen_update(n)
        sys.exit(0)
    if sys.argv[1:2] == ["--keys-only"]:
        for n in MODELS:
            gen_keys(n)
        sys.exit(0)
    if sys.argv[1:2] == ["--transform-only"]:
        for n in TRANSFORM_MODELS:
            gen_transform(n)
        sys.exit(0)
    only = sys.argv[1:]
    if not only:
        gen_connectivity()
    for n in MODELS:
        if not only or n in only:
            gen_forward(n)
            gen_keys(n)
            gen_update(n)
    if not only:
        gen_train()
        for n in TRANSFORM_MODELS:
            gen_transform(n)
