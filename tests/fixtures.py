"""Deterministic parameters / statistics shared by the golden generator, the tests and the bench.

Parameters are a pure function of their state_dict key (numpy RandomState seeded by crc32 of the
key), so the reference model, the oracle and the CUDA modules get bit-identical weights without
shipping an 8.8 MB checkpoint per model and without depending on module construction order.
Distributions follow PyTorch's defaults: Linear weight and bias
U(+-1/sqrt(fan_in)); LayerNorm affine is perturbed away from (1, 0) so that it is exercised.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

STAT_KEYS = [
    "cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
    "cell_pressure", "face_velocity_difference_x", "face_velocity_difference_y",
    "face_edge_vector_x", "face_edge_vector_y", "face_area", "face_velocity_x", "face_velocity_y",
    "face_pressure", "face_flux", "face_adjacent_distance", "face_velocity_diff_char",
    "cell_velocity_char",
]


def default_stats():
    """Non-trivial per-key statistics (mean/std/min/max) so (de)normalisation is exercised."""
    out = {}
    for i, k in enumerate(STAT_KEYS):
        out[k] = {"mean": 0.05 * (i % 5) + 0.1, "std": 1.0 + 0.1 * (i % 4), "min": -1.0, "max": 1.0}
    return out


def stats_for(model_name: str):
    """``default_stats()`` plus the extra registry keys a model family uses (ConservativeH: the std-scaled
    face_velocity_diff_x / _y, Conservative.py:948-962).  Kept separate so the other models' fixtures do not move."""
    out = default_stats()
    if model_name in ("ConservativeH", "ConservativeJ", "ConservativeK"):
        for i, k in enumerate(("face_velocity_diff_x", "face_velocity_diff_y")):
            out[k] = {"mean": 0.07 + 0.03 * i, "std": 1.2 + 0.1 * i, "min": -1.0, "max": 1.0}
    if model_name == "FvgnE":      # physical normalisation (Fvgn.py:846-851)
        for i, k in enumerate(("characteristic_velocity", "characteristic_length", "characteristic_pressure")):
            out[k] = {"mean": 0.6 + 0.2 * i, "std": 1.1, "min": 0.1, "max": 1.5 + 0.5 * i}
    if model_name == "FvgnH":      # augmented face features (Fvgn.py:1066-1081)
        for i, k in enumerate(("face_normal_x", "face_normal_y", "face_angle")):
            out[k] = {"mean": 0.04 * (i + 1), "std": 0.9 + 0.1 * i, "min": -1.0, "max": 1.0}
    return out


def add_mls_fixture(c_graph, k: int = 8, seed: int = 11):
    """Synthetic moving-least-squares stencil on a cell graph: ``grad_neighbours`` [N, k] (random cells) and
    ``grad_weights`` [N, k, 2].  The reference computes these offline (utils/maths.py MovingLeastSquaresWeights);
    the models only consume them (StreamFunc.py:100-105, fvm.py:40-52), so any values exercise the path."""
    g = torch.Generator().manual_seed(seed)
    n = c_graph.pos.shape[0]
    c_graph.grad_neighbours = torch.randint(0, n, (n, k), generator=g)
    c_graph.grad_weights = torch.randn(n, k, 2, generator=g) * 0.3
    return c_graph


def raw_graphs(mesh, seed: int = 17, steps: int = 4):
    """Raw (pre-``transform_features``) graph triplet on a synthetic mesh: velocity / pressure / flux time series on
    cells and faces plus the static geometry, in the shapes the reference's dataset hands to the model class
    (``src/datasets/DataSet.py:210-274``: series are [rows, time, channels], face type is [E, 1])."""
    from gnn_fluid_dynamics_b200.graph import Data
    g = torch.Generator().manual_seed(seed)
    n, e = mesh.n_cells, mesh.n_faces
    f32 = torch.float32
    c = Data(velocity=torch.randn(n, steps, 2, generator=g), pressure=torch.randn(n, steps, 1, generator=g),
             pos=torch.from_numpy(mesh.cell_pos).to(f32), volume=torch.from_numpy(mesh.cell_volume).to(f32),
             normal=torch.from_numpy(mesh.cell_normal).to(f32),
             edge_index=torch.from_numpy(mesh.cell_edge_index.copy()), dt=torch.tensor(0.01, dtype=f32))
    f = Data(velocity=torch.randn(e, steps, 2, generator=g), pressure=torch.randn(e, steps, 1, generator=g),
             flux=torch.randn(e, steps, 1, generator=g), pos=torch.from_numpy(mesh.face_pos).to(f32),
             face=torch.from_numpy(mesh.face_index.copy()), type=torch.from_numpy(mesh.face_type.copy()),
             area=torch.from_numpy(mesh.face_area).to(f32), normal=torch.from_numpy(mesh.face_normal).to(f32))
    v = Data(pos=torch.from_numpy(mesh.vertex_pos).to(f32), edge_index=torch.from_numpy(mesh.vertex_edge_index.copy()),
             face=torch.from_numpy(mesh.cells.T.copy()))
    return [c, f, v]


def _rs(key: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


@torch.no_grad()
def fill_state_dict_deterministic(module: torch.nn.Module, seed: int = 1) -> None:
    """Overwrite every Linear / LayerNorm / BatchNorm tensor of ``module`` in place."""
    fill_tensor_dict(module.state_dict(), seed)


@torch.no_grad()
def state_dict_from_keys(model_name: str, seed: int = 1):
    """The deterministic state_dict of ``model_name`` WITHOUT constructing the model: key names and shapes from the
    reference-generated ``tests/golden/keys_{model}.json``, normaliser buffers from ``stats_for``, everything else from
    the same key-seeded fill.  Used by the CPU reference arm of bench.py, which must not load the CUDA library."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    keys = json.load(open(os.path.join(here, "golden", f"keys_{model_name}.json")))
    stats = stats_for(model_name)
    sd = {}
    for key, shape in keys:
        if key.startswith("normalizer."):
            name, stat = key[len("normalizer."):].rsplit("_", 1)
            sd[key] = torch.tensor(stats[name][stat], dtype=torch.float)
        elif key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros(shape, dtype=torch.long)
        else:
            sd[key] = torch.zeros(shape, dtype=torch.float32)
            if key.endswith("anisotropy_ratio"):
                sd[key].fill_(0.0001)
    fill_tensor_dict(sd, seed)
    return sd


@torch.no_grad()
def fill_tensor_dict(sd, seed: int = 1) -> None:
    for key, t in sd.items():
        if "normalizer." in key:
            continue
        rs = _rs(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        base = key.rsplit(".", 1)[0]
        if leaf == "weight" and t.dim() == 2:
            bound = 1.0 / np.sqrt(t.shape[1])
            t.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(t.shape)).astype(np.float32)))
        elif leaf == "bias" and (base + ".weight") in sd and sd[base + ".weight"].dim() == 2:
            bound = 1.0 / np.sqrt(sd[base + ".weight"].shape[1])
            t.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(t.shape)).astype(np.float32)))
        elif leaf == "weight":          # LayerNorm / BatchNorm scale
            t.copy_(torch.from_numpy((1.0 + 0.1 * rs.uniform(-1, 1, size=tuple(t.shape))).astype(np.float32)))
        elif leaf == "bias":
            t.copy_(torch.from_numpy((0.1 * rs.uniform(-1, 1, size=tuple(t.shape))).astype(np.float32)))
        elif leaf == "running_mean":
            t.fill_(0.2)
        elif leaf == "running_var":
            t.fill_(1.5)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / ||b||_2 (the parity metric of BASELINE.json: north_star)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
