"""Shared test helpers: configs, deterministic models, fixtures."""
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_W = {"continuity": 0, "cell_velocity_change": 10, "cell_pressure": 1, "face_velocity": 1,
          "face_flux": 1, "face_pressure": 1, "cell_velocity": 10}
LOSS_MODELS = ("FvgnC", "VertPotC", "VertPotE", "VertPotG", "ConservativeB", "ConservativeJ", "FvgnB", "FvgnE", "FvgnH", "FvgnJ", "FvgnK", "MgnB", "MgnC", "StreamFuncA", "StreamFuncB", "StreamFuncC", "StreamFuncD", "FluxB", "FluxC", "FluxD")
# (mesh kind, feature flavour) used by tests/golden/make_golden.py per model
GOLDEN_SETUP = {"MgnA": ("cylinder", "fvgn"), "FvgnA": ("cylinder", "fvgn"), "FluxA": ("ellipse", "fvgn"),
                "ConservativeA": ("cylinder", "conservative"), "VertPotA": ("airfoil", "fvgn"),
                "ConservativeE": ("ellipse", "fvgn"), "ConservativeF": ("airfoil", "fvgn"),
                "ConservativeD": ("ellipse", "conservative"), "ConservativeG": ("cylinder", "fvgn"),
                "ConservativeI": ("airfoil", "fvgn"), "ConservativeH": ("cylinder", "conservative_h"),
                "FvgnF": ("airfoil", "fvgn"), "ConservativeK": ("ellipse", "conservative_h"),
                "MgnB": ("ellipse", "fvgn"), "MgnC": ("airfoil", "fvgn"), "StreamFuncA": ("cylinder", "fvgn"),
                "StreamFuncB": ("ellipse", "fvgn"), "StreamFuncC": ("airfoil", "fvgn"), "StreamFuncD": ("cylinder", "fvgn"),
                "FluxB": ("cylinder", "fvgn"), "FluxC": ("airfoil", "fvgn"), "FluxD": ("ellipse", "fvgn"),
                "ConservativeB": ("airfoil", "conservative"), "ConservativeJ": ("ellipse", "conservative_h"),
                "VertPotB": ("cylinder", "fvgn"), "VertPotC": ("ellipse", "fvgn"), "VertPotE": ("airfoil", "fvgn"), "VertPotG": ("cylinder", "fvgn"),
                "FvgnC": ("ellipse", "fvgn"), "FvgnB": ("cylinder", "fvgn"), "FvgnD": ("ellipse", "fvgn"), "FvgnE": ("airfoil", "fvgn"), "FvgnH": ("cylinder", "fvgn"), "FvgnI": ("ellipse", "fvgn"), "FvgnJ": ("airfoil", "fvgn"), "FvgnK": ("cylinder", "fvgn")}
FVGN_LIKE = ("FvgnA", "FvgnB", "FvgnD", "FvgnE", "FvgnH", "FvgnI", "FvgnJ", "FvgnK")
MGN_LIKE = ("MgnA", "MgnB", "MgnC", "StreamFuncA", "StreamFuncB", "StreamFuncC", "StreamFuncD")
ALL_MODELS = list(GOLDEN_SETUP)


def make_config(mp_num=15, precision=None):
    return NS(model=NS(hidden_width=128, mp_num=mp_num, precision=precision, bundle_size=3),
              training=NS(dropout_rate=0.0, loss_weights=dict(LOSS_W)))


def mse(output, target, mask, batch=None):
    if mask is not None:
        output, target = output[mask], target[mask]
    return torch.nn.functional.mse_loss(output, target)


def build_model(name, mp_num=15, precision=None, seed=1, device="cpu"):
    from gnn_fluid_dynamics_b200.models import MODEL_CLASSES
    from fixtures import fill_state_dict_deterministic, stats_for
    model = MODEL_CLASSES[name](make_config(mp_num, precision), mse, None, stats_for(name))
    fill_state_dict_deterministic(model, seed=seed)
    return model.to(device)


def golden_graphs(name, flip=False, n_cells=160, mesh_seed=3, feat_seed=5):
    """Same construction as tests/golden/make_golden.py:graphs_for."""
    from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs
    kind, flavour = GOLDEN_SETUP[name]
    mesh = make_mesh(n_cells, kind, seed=mesh_seed)
    g = mesh_graphs(mesh, seed=feat_seed, flavour=flavour, flip_edges=flip)
    return mesh, finish_graphs(name, g)


def finish_graphs(name, g):
    """Model-specific targets / extra inputs on top of ``mesh_graphs`` (shared with bench.py's synthetic batches)."""
    c, f, v = g
    if name in MGN_LIKE + ("ConservativeB",):
        c.y = torch.cat([c.y, torch.randn(c.x.shape[0], 1, generator=torch.Generator().manual_seed(9))], 1)
        f.y = f.y[:, :2].contiguous()
    elif name in FVGN_LIKE + ("ConservativeA", "ConservativeE", "ConservativeF", "ConservativeD", "ConservativeG", "ConservativeI", "ConservativeH", "FvgnF", "ConservativeK", "ConservativeJ"):
        f.y = f.y[:, :3].contiguous()
    if name == "FluxC":
        f.y = f.y[:, :2].contiguous()     # (pressure, flux) targets, Flux.py:322
    fvgn_variant_fixture(name, c, f)
    if name == "ConservativeI":
        f.type = f.type.reshape(-1)      # see tests/golden/make_golden.py: the reference needs a 1-D type tensor here
    if name.startswith("StreamFunc") or name in ("MgnB", "MgnC"):
        from fixtures import add_mls_fixture
        add_mls_fixture(c)
    c.batch = torch.zeros(c.x.shape[0], dtype=torch.long)
    f.batch = torch.zeros(f.pos.shape[0], dtype=torch.long)
    return g


def fvgn_variant_fixture(name, c, f):
    """Extra inputs of the FvgnA glue variants (same construction in tests/golden/make_golden.py)."""
    from fixtures import add_mls_fixture
    if name in ("VertPotC", "VertPotE"):      # FluxC targets: (pressure, flux)
        f.y = f.y[:, :2].contiguous()
    if name in ("FvgnB", "VertPotB"):       # face moving-least-squares stencil for the diffusion term (Fvgn.py:446)
        add_mls_fixture(f, seed=12)
    if name == "FvgnH":       # 7 + 5 face feature columns (Fvgn.py:1057)
        extra = torch.randn(f.x.shape[0], 2, generator=torch.Generator().manual_seed(13))
        f.x = torch.cat([f.x[:, :5], extra, f.x[:, 5:]], dim=1)
    if name == "FvgnC":       # temporal bundle of 3 target steps (Fvgn.py:484, 506)
        c.y = torch.randn(c.x.shape[0], 3, 2, generator=torch.Generator().manual_seed(21))
        f.y = torch.randn(f.x.shape[0], 3, 3, generator=torch.Generator().manual_seed(22))
    if name == "FvgnK":       # per-mesh Reynolds number; the reference needs a 1-D type tensor here (Fvgn.py:1291-1296)
        c.Re = torch.tensor([150.0])
        f.type = f.type.reshape(-1)


def load_golden(fname):
    return {k: v for k, v in np.load(os.path.join(GOLDEN, fname), allow_pickle=False).items()}


def graphs_to(graphs, device):
    return [g.to(device) for g in graphs]
