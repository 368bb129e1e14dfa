"""Flux family on the B200 kernels - drop-in for reference ``src/models/Flux.py`` (FluxA).

The encoder / processor / decoder are FvgnA's unchanged (Flux.py:26-33); FluxA adds a face-flux
output channel (decoder width 6) and its own integrator (Flux.py:157-206).
"""
from __future__ import annotations

import torch
from torch import nn

from ..topology import get_topology
from .base import col, n_class_types
from .Fvgn import FvgnA, normalize_face_area


def normalize_vol_dt(cell_volume, edge_index, dt, batch_norm):   # utils/normalisation.py:346-365
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
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        cell_flux = face_flux_to_cell_flux(output[1][:, 3:4], f_graph.face, c_graph.edge_index)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, 0:2],
                "face_pressure": output[1][:, 2:3], "face_flux": output[1][:, 3:4],
                "cell_flux": cell_flux.squeeze(-1)}

    class Integrator(nn.Module):   # Flux.py:157-206
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.vol_dt_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
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
