"""Oracle: whole-model forward for the two reference drivers the benches use (FvgnA, MgnA):
normalise -> encoder -> blocks -> decoder -> integrator -> (de)normalise, CPU fp32.

Follows Fvgn.py:150-174 + Integrator Fvgn.py:214-255 and Mgn.py:153-173; normalisation follows
utils/normalisation.py:255-290 (z_score only: every key those two models register is 'z_score').
Graph objects are anything with attribute access (``gnn_fluid_dynamics_b200.graph.Data``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .blocks import processor_fwd

# (graph index, attribute, column, stats key)  - Fvgn.py:72-88 / Mgn.py:112-128
_FVGN_INPUTS = [
    (0, "x", 0, "cell_velocity_x"), (0, "x", 1, "cell_velocity_y"),
    (0, "y", 0, "cell_velocity_change_x"), (0, "y", 1, "cell_velocity_change_y"),
    (1, "x", 0, "face_velocity_difference_x"), (1, "x", 1, "face_velocity_difference_y"),
    (1, "x", 2, "face_edge_vector_x"), (1, "x", 3, "face_edge_vector_y"), (1, "x", 4, "face_area"),
    (1, "y", 0, "face_velocity_x"), (1, "y", 1, "face_velocity_y"), (1, "y", 2, "face_pressure"),
]
_FVGN_OUTPUTS = [(0, 0, "cell_velocity_change_x"), (0, 1, "cell_velocity_change_y"),
                 (1, 0, "face_velocity_x"), (1, 1, "face_velocity_y"), (1, 2, "face_pressure")]
_MGN_INPUTS = [
    (0, "x", 0, "cell_velocity_x"), (0, "x", 1, "cell_velocity_y"),
    (1, "x", 0, "face_velocity_difference_x"), (1, "x", 1, "face_velocity_difference_y"),
    (1, "x", 2, "face_edge_vector_x"), (1, "x", 3, "face_edge_vector_y"), (1, "x", 4, "face_area"),
    (0, "y", 0, "cell_velocity_change_x"), (0, "y", 1, "cell_velocity_change_y"),
    (0, "y", 2, "cell_pressure"),
    (1, "y", 0, "cell_velocity_x"), (1, "y", 1, "cell_velocity_y"),
]
_MGN_OUTPUTS = [(0, 0, "cell_velocity_change_x"), (0, 1, "cell_velocity_change_y"),
                (0, 2, "cell_pressure")]


def _z(data, stats, key, inverse=False):
    mean = torch.tensor(stats[key]["mean"], dtype=torch.float)
    std = torch.tensor(stats[key]["std"], dtype=torch.float)
    std = torch.clamp(std, min=1e-8) + 1e-8
    return data * std + mean if inverse else (data - mean) / std


def normalise_inputs(model_name, stats, graphs):
    """In place, like the reference (normalisation.py:243, 262)."""
    table = _FVGN_INPUTS if model_name in ("FvgnA", "FluxA") else _MGN_INPUTS
    for gi, attr, col, key in table:
        t = getattr(graphs[gi], attr)
        if t.shape[1] > col:
            t[:, col:col + 1] = _z(t[:, col:col + 1], stats, key)
    return graphs


def topo_of(graphs):
    c, f, v = graphs
    return {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face,
            "n_vertices": v.num_nodes}


def fvgn_integrator(sd, edge_out, c, f, training=False):
    """Integrator.forward (Fvgn.py:221-255) with normalize_face_area (normalisation.py:325-344).
    BatchNorm1d(1) in eval mode uses running stats; in train mode batch stats."""
    dt_mean = torch.mean(c.dt)
    vol = (c.volume.index_select(0, c.edge_index[0]) + c.volume.index_select(0, c.edge_index[1])) / 2
    raw = (f.area * (dt_mean / vol)).view(-1, 1)
    p = "integrator.face_area_norm."
    area = F.batch_norm(raw, sd[p + "running_mean"].detach().clone(), sd[p + "running_var"].detach().clone(),
                        sd[p + "weight"], sd[p + "bias"], training=training, momentum=0.1, eps=1e-5)
    cf = f.face
    unv = c.normal
    uv, pf, flux_d = edge_out[:, :2], edge_out[:, 2:3], edge_out[:, 3:]
    uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)

    def dot2(a, n):  # chain_flux_dot_product (maths.py:12-20)
        return torch.cat([(a[:, 0:2] * n).sum(-1, keepdim=True), (a[:, 2:4] * n).sum(-1, keepdim=True)], -1)

    phi_a = sum(dot2(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
    phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
    phi_p = sum(pf[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
    return 1.0 * (-phi_a - phi_p / 1) + phi_d


def model_forward(model_name, sd, stats, graphs, mp_num, mode="train", training=False):
    """Whole forward of FvgnA / MgnA -> (dict of outputs, dict of processor intermediates)."""
    graphs = normalise_inputs(model_name, stats, graphs)
    c, f, v = graphs
    if model_name == "FvgnA":
        mid = processor_fwd("fvgn", sd, c.x, f.x, topo_of(graphs), mp_num)
        edge_out = mid["dec"]
        acc = fvgn_integrator(sd, edge_out, c, f, training=training)
        out = [acc, edge_out]
        if mode == "rollout":
            out = [o.clone() for o in out]
            for oi, col, key in _FVGN_OUTPUTS:
                out[oi][:, col:col + 1] = _z(out[oi][:, col:col + 1], stats, key, inverse=True)
        return {"cell_velocity_change": out[0][:, 0:2], "face_velocity": out[1][:, :2],
                "face_pressure": out[1][:, 2:3]}, mid
    if model_name == "MgnA":
        mid = processor_fwd("mgn", sd, c.x, f.x, topo_of(graphs), mp_num)
        out = [mid["dec"]]
        if mode == "rollout":
            out = [o.clone() for o in out]
            for oi, col, key in _MGN_OUTPUTS:
                out[oi][:, col:col + 1] = _z(out[oi][:, col:col + 1], stats, key, inverse=True)
        return {"cell_velocity_change": out[0][:, 0:2], "cell_pressure": out[0][:, 2:3]}, mid
    raise ValueError(model_name)


def fvgn_loss(sd, out, graphs, loss_weights, training=True):
    """FvgnA.loss (Fvgn.py:176-212) with MSE_per_element_torch (utils/loss.py:55-60)."""
    c, f, v = graphs
    mse = lambda a, b: torch.mean((a - b) ** 2)
    dt_mean = torch.mean(c.dt)
    vol = (c.volume.index_select(0, c.edge_index[0]) + c.volume.index_select(0, c.edge_index[1])) / 2
    p = "integrator.face_area_norm."
    area = F.batch_norm((f.area * (dt_mean / vol)).view(-1, 1), sd[p + "running_mean"].detach().clone(),
                        sd[p + "running_var"].detach().clone(), sd[p + "weight"], sd[p + "bias"],
                        training=training, momentum=0.1, eps=1e-5)
    ff, unv, fv = f.face, c.normal, out["face_velocity"]
    div = sum((fv[ff[j]] * unv[:, j, :]).sum(-1, keepdim=True) * area[ff[j]] for j in range(3))
    cont = mse(div, torch.zeros_like(div))
    cvc = mse(out["cell_velocity_change"], c.y)
    interior = ~f.boundary_mask
    fvl = mse(out["face_velocity"][interior], f.y[:, :2][interior])
    fpl = mse(out["face_pressure"], f.y[:, 2:3])
    w = loss_weights
    total = (w["continuity"] * cont + w["cell_velocity_change"] * cvc + w["face_velocity"] * fvl
             + w["face_pressure"] * fpl)
    return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": cont,
            "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}


def rollout_step(model_name, sd, stats, graphs, mp_num):
    """One autoregressive step on CPU: forward in 'rollout' mode on clones (rollout.py:313), velocity update
    (rollout.py:340) and update_features (Fvgn.py:133-148 / Mgn.py:139-151), in place on ``graphs``."""
    c, f, v = graphs
    out, _ = model_forward(model_name, sd, stats, [g.clone() for g in graphs], mp_num, mode="rollout")
    vel = c.x[:, :2] + out["cell_velocity_change"]
    c.x = vel.detach()
    u = c.x[:, :2]
    dv = u[c.edge_index[0]] - u[c.edge_index[1]]
    if model_name == "MgnA":
        mask = f.boundary_mask
    else:
        mask = ((f.type == 2) | (f.type == 1)).squeeze(-1)      # INFLOW | WALL_BOUNDARY (OpenFoam.py:19-24)
    dv[mask] = f.y[:, 0:2][mask]
    f.x[:, 0:2] = dv
    return vel
