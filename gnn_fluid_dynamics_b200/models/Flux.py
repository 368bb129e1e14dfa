"""Flux family on the B200 kernels - drop-in for reference ``src/models/Flux.py`` (FluxA, FluxB, FluxC, FluxD).

The encoder / processor / decoder are FvgnA's unchanged (Flux.py:26-33); FluxA adds a face-flux
output channel (decoder width 6) and its own integrator (Flux.py:157-206).
"""
from __future__ import annotations

import torch
from torch import nn

from ..topology import get_topology
from .base import col, n_class_types
from .Fvgn import FvgnA, graph_topology, normalize_face_area

import os
USE_GATHER3 = os.environ.get("GNNFD_GATHER3", "1") != "0"      # A/B knob


def normalize_vol_dt(cell_volume, edge_index, dt, batch_norm, topo=None):   # utils/normalisation.py:346-365
    """``BatchNorm1d(1)(1.0 * mean(dt) / mean adjacent cell volume)``.  With the mesh topology at hand this is
    ``normalize_face_area`` of a unit area: the same fused kernel pair (batch statistics, running-stat update, normalisation;
    one backward kernel) - ATen's single-channel BatchNorm backward reduces 242k rows in ONE block (6.7 ms on the 8 x 20k
    batch when its incoming gradient is a strided view)."""
    if topo is not None and cell_volume.is_cuda and isinstance(batch_norm, nn.BatchNorm1d):
        from ..fvm_ops import face_area_norm
        ones = torch.ones(edge_index.shape[1], dtype=torch.float32, device=cell_volume.device)
        return face_area_norm(ones, cell_volume, topo.row, topo.col, dt, batch_norm)
    vol = (cell_volume.index_select(0, edge_index[0]) + cell_volume.index_select(0, edge_index[1])) / 2
    return batch_norm((1.0 * (torch.mean(dt) / vol)).view(-1, 1))


def face_flux_to_cell_flux(face_flux, face_face, cell_adjacency):
    """Owner-oriented face flux -> signed per-cell local face flux [N,3,1] (utils/fvm.py:96-156):
    +1 for the owner, -1 for the neighbour of an interior face, boundary faces keep the owner sign."""
    n = face_face.shape[1]
    fidx = face_face.t().reshape(-1)
    cidx = torch.arange(n, device=face_flux.device).repeat_interleave(3)
    owners, neigh = cell_adjacency[0, fidx], cell_adjacency[1, fidx]
    interior = ~((owners == neigh) | (neigh == -1))
    signs = torch.zeros_like(cidx)
    signs = torch.where(cidx == owners, torch.ones_like(signs), signs)
    signs = torch.where(interior & (cidx == neigh), -torch.ones_like(signs), signs)
    return (face_flux.view(-1)[fidx] * signs).view(n, 3).unsqueeze(-1)


class FluxA(FvgnA):
    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 6, 0])   # Flux.py:35-37

    @classmethod
    def normalisation_tables(cls):   # Flux.py:39-55
        kinds, inputs, outputs = super().normalisation_tables()
        kinds["face_flux"] = "z_score"
        inputs = inputs + [(1, "y", col(3), "face_flux")]
        outputs = outputs + [(1, col(3), "face_flux")]
        return kinds, inputs, outputs

    def forward(self, graphs, mode="rollout"):   # Flux.py:89-116
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        c_graph.topology = topo          # the integrator reuses its int32 index tensors
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        if output[1].is_cuda and not (torch.is_grad_enabled() and output[1].requires_grad):
            from ..fvm_ops import cell_faces, flux_integrate      # one kernel instead of ~15 index / where kernels
            cell_flux = flux_integrate(output[1], None, None, None, cell_faces(topo, f_graph.face), topo.row, topo.col,
                                       want_acc=False, want_cell_flux=True, flux_col=3)
        else:
            cell_flux = face_flux_to_cell_flux(output[1][:, 3:4], f_graph.face, c_graph.edge_index).squeeze(-1)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, 0:2],
                "face_pressure": output[1][:, 2:3], "face_flux": output[1][:, 3:4],
                "cell_flux": cell_flux}

    def loss(self, output, graphs):   # Flux.py:118-155
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        div = output["cell_flux"][:, 0] + output["cell_flux"][:, 1] + output["cell_flux"][:, 2]   # fvm.py:13-19
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fvl = lf(output["face_velocity"], f_graph.y[:, :2], ~f_graph.boundary_mask, f_graph.batch)
        ffl = lf(output["face_flux"], f_graph.y[:, 3:4], None, f_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc + w["face_velocity"] * fvl
                 + w["face_flux"] * ffl + w["face_pressure"] * fpl)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_flux_loss": ffl,
                "face_pressure_loss": fpl}

    class Integrator(nn.Module):   # Flux.py:157-206
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.vol_dt_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
            topo = graph_topology(c_graph)
            if (topo is not None and edge_output.is_cuda and edge_output.shape[1] == 6
                    and not (torch.is_grad_enabled() and edge_output.requires_grad)):
                # evaluation / rollout: signs, gathers, products and sums of the expression below as one kernel
                # (bit-identical: every operation separately rounded in the same order)
                from ..fvm_ops import cell_faces, flux_integrate
                coeff = normalize_vol_dt(c_graph.volume, c_graph.edge_index, dt, self.vol_dt_norm, topo=topo)
                area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm, topo=topo)
                self.face_area = area
                return flux_integrate(edge_output, coeff, area, unv, cell_faces(topo, cf), topo.row, topo.col, self.rho)
            if topo is not None and edge_output.is_cuda and edge_output.shape[1] == 6 and USE_GATHER3:
                # training: the same tensor expression with its gathers as two gather3 launches (sort-free autograd)
                from ..fvm_ops import cell_faces, gather3
                cfs = cell_faces(topo, cf)
                cell_flux = face_flux_to_cell_flux(edge_output[:, 3:4], cf, c_graph.edge_index)
                coeff = normalize_vol_dt(c_graph.volume, c_graph.edge_index, dt, self.vol_dt_norm, topo=topo)
                area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm, topo=topo)
                self.face_area = area
                eo = gather3(edge_output, cfs, topo.row, topo.col)                              # [3, N, (u, v, p, phi, d0, d1)]
                ka = gather3(torch.cat([coeff, area], dim=1), cfs, topo.row, topo.col)          # [3, N, (coeff, area)]
                phi_a = sum(eo[j][:, 0:2] * cell_flux[:, j] * ka[j][:, 0:1] for j in range(3))
                phi_d = eo[0][:, 4:6] + eo[1][:, 4:6] + eo[2][:, 4:6]
                phi_p = sum(eo[j][:, 2:3] * unv[:, j, :] * ka[j][:, 1:2] for j in range(3))
                return 1.0 * (-phi_a - phi_p / self.rho) + phi_d
            uv, p_face = edge_output[:, :2], edge_output[:, 2:3]
            flux_face, flux_d = edge_output[:, 3:4], edge_output[:, 4:6]
            cell_flux = face_flux_to_cell_flux(flux_face, cf, c_graph.edge_index)
            coeff = normalize_vol_dt(c_graph.volume, c_graph.edge_index, dt, self.vol_dt_norm)
            phi_a = sum(uv[cf[j]] * cell_flux[:, j] * coeff[cf[j]] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d


def _face_losses(model, output, graphs, flux_col, pressure_col):
    """The flux-divergence loss shared by FluxB (Flux.py:250-283) and FluxC (Flux.py:423-456)."""
    c_graph, f_graph, v_graph = graphs
    lf = model.mse_term
    ff, flux = f_graph.face, output["face_flux"]
    div = flux[ff[0]] + flux[ff[1]] + flux[ff[2]]                     # fvm.divergence_from_face_flux
    continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
    cvc = lf(output["cell_velocity_change"], c_graph.y[:, 0:2], None, c_graph.batch)
    ffl = lf(flux, f_graph.y[:, flux_col:flux_col + 1], None, f_graph.batch)
    fpl = lf(output["face_pressure"], f_graph.y[:, pressure_col:pressure_col + 1], None, f_graph.batch)
    w = model.config.training.loss_weights
    total = w["continuity"] * continuity + w["cell_velocity_change"] * cvc + w["face_flux"] * ffl + w["face_pressure"] * fpl
    return total, continuity, cvc, ffl, fpl


class FluxB(FluxA):
    """Predicts the face velocity only; the flux is u_f . n A and enters the loss (Flux.py:209-283).  FvgnA's
    integrator, decoder width 5."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = FvgnA.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 5, 0])

    def forward(self, graphs, mode="rollout"):   # Flux.py:219-248
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        face_area = self.integrator.face_area
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
            face_area = f_graph.area
        face_flux = (torch.sum(output[1][:, 0:2] * f_graph.normal, dim=-1, keepdim=True) * face_area).reshape(-1, 1)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, 0:2],
                "face_pressure": output[1][:, 2:3], "face_flux": face_flux}

    def loss(self, output, graphs):
        total, continuity, cvc, ffl, fpl = _face_losses(self, output, graphs, 3, 2)
        return {"total_log_loss": torch.mean(torch.log(total)), "cell_velocity_change_loss": cvc,
                "face_flux_loss": ffl, "face_pressure_loss": fpl}


def cell_to_face(cell_values, cell_edge_index, face_centre, cell_centres):   # utils/geometry.py:460-491
    i0, i1 = cell_edge_index[0], cell_edge_index[1]
    w0 = 1.0 / (torch.norm(face_centre - cell_centres[i0], dim=1) + 1e-10)
    w1 = 1.0 / (torch.norm(face_centre - cell_centres[i1], dim=1) + 1e-10)
    w1 = torch.where(i0 == i1, torch.zeros_like(w1), w1)
    s = w0 + w1
    return (w0 / s)[:, None] * cell_values[i0] + (w1 / s)[:, None] * cell_values[i1]


class FluxC(FvgnA):
    """Predicts pressure, flux and diffusion only; the advected face velocity is interpolated from the cells
    (Flux.py:286-456).  Decoder width 4."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 4, 0])

    @classmethod
    def normalisation_tables(cls):   # Flux.py:326-354
        kinds, inputs, outputs = FvgnA.normalisation_tables()
        drop = ("face_velocity_x", "face_velocity_y", "face_pressure")
        kinds = {k: v for k, v in kinds.items() if k not in drop}
        kinds.update({"face_pressure": "z_score", "face_flux": "z_score"})
        inputs = [r for r in inputs if not (r[0] == 1 and r[1] == "y")]
        inputs += [(1, "y", col(0), "face_pressure"), (1, "y", col(1), "face_flux")]
        outputs = [r for r in outputs if r[0] != 1] + [(1, col(0), "face_pressure"), (1, col(1), "face_flux")]
        return kinds, inputs, outputs

    def forward(self, graphs, mode="rollout"):   # Flux.py:356-380
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "face_pressure": output[1][:, 0:1],
                "face_flux": output[1][:, 1:2]}

    def loss(self, output, graphs):
        total, continuity, cvc, ffl, fpl = _face_losses(self, output, graphs, 1, 0)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_flux_loss": ffl, "face_pressure_loss": fpl}

    class Integrator(nn.Module):   # Flux.py:382-421
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
            uv = cell_to_face(c_graph.x[:, 0:2], c_graph.edge_index, f_graph.pos, c_graph.pos)
            p_face, flux_face, flux_d = edge_output[:, 0:1], edge_output[:, 1:2], edge_output[:, 2:4]
            phi_a = sum(uv[cf[j]] * flux_face[cf[j]] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d


class FluxD(FluxA):
    """FluxA's network with learnt output scales and a physical (un-normalised) integrator (Flux.py:459-595)."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1.0)
        self.velocity_scale_x = nn.Parameter(torch.tensor(0.1))
        self.velocity_scale_y = nn.Parameter(torch.tensor(0.0001))
        self.pressure_scale = nn.Parameter(torch.tensor(0.01))
        self.diffusion_scale = nn.Parameter(torch.tensor(0.01))
        self.flux_scale = nn.Parameter(torch.tensor(0.001))

    def forward(self, graphs, mode="rollout"):   # Flux.py:477-515
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, raw = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        edge_attr_out = torch.cat([raw[:, 0:1] * self.velocity_scale_x, raw[:, 1:2] * self.velocity_scale_y,
                                   raw[:, 2:3] * self.pressure_scale, raw[:, 3:4] * self.flux_scale,
                                   raw[:, 4:6] * self.diffusion_scale], dim=-1)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode != "rollout":
            output = self.normalizer.output(output)      # normalised for the training loss
        cell_flux = face_flux_to_cell_flux(output[1][:, 3:4], f_graph.face, c_graph.edge_index)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, 0:2],
                "face_pressure": output[1][:, 2:3], "face_flux": output[1][:, 3:4],
                "cell_flux": cell_flux.squeeze(-1)}

    class Integrator(nn.Module):   # Flux.py:556-595
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.nu = 0.001

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf, area = c_graph.normal, f_graph.face, f_graph.area
            uv, p_face = edge_output[:, :2], edge_output[:, 2:3]
            flux_face, flux_d = edge_output[:, 3:4], edge_output[:, 4:6]
            cell_flux = face_flux_to_cell_flux(flux_face, cf, c_graph.edge_index)
            phi_a = sum(uv[cf[j]] * cell_flux[:, j] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p / self.rho + self.nu * phi_d)
