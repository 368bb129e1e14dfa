"""VertPot family on the B200 kernels - drop-in for reference ``src/models/VertPot.py`` (VertPotA):
FvgnA blocks (under ``node_block`` / ``edge_block``) plus a Vertex_Block that sums the face block's raw
output onto vertices; edge + vertex decoder heads (VertPot.py:187-231).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import processor as P
from ..topology import get_topology
from .base import build_mlp, col, n_class_types
from .Flux import FluxA, normalize_vol_dt
from .Fvgn import FvgnA, normalize_face_area


def cell_flux_from_vertices(vertex_out, v_face):
    """Per-cell edge differences of the vertex potential (VertPot.py:24-40): [N,3]."""
    v = vertex_out[v_face]
    return torch.stack([v[1] - v[2], v[2] - v[0], v[0] - v[1]], dim=0).squeeze(-1).T


class VertPotA(FluxA):
    family = "vertpot"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)
        self.integrator = self.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 5, 1])   # VertPot.py:59-61

    @classmethod
    def normalisation_tables(cls):   # VertPot.py:63-72
        kinds, inputs, outputs = super().normalisation_tables()
        return kinds, inputs, outputs + [(0, col(2, 5), "face_flux")]

    def training_plan(self):
        from ..training import Plan, Site
        plan = getattr(self, "_gnnfd_plan", None)
        if plan is None:
            blocks = [(Site(b.node_block.cell_mlp), Site(b.edge_block.face_mlp)) for b in self.processer_list]
            plan = Plan("vertpot", Site(self.encoder.face_mlp), Site(self.encoder.cell_mlp), blocks,
                        Site(self.decoder.edge_mlp), dec_vertex=Site(self.decoder.vertex_mlp))
            object.__setattr__(self, "_gnnfd_plan", plan)
        return plan

    def encode_process_decode(self, c_x, f_x, topo, hook=None):
        prec = self.prec
        if self.wants_grad():   # training step: kernel-scheduled backward (training.py), both decoder heads
            from ..training import encode_process_decode_train
            edge_out, vertex_out = encode_process_decode_train(self.training_plan(), topo, prec, c_x, f_x)
            return None, None, None, edge_out, vertex_out
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        x, e, vx = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook)
        edge_out = P.mlp_rows(self.decoder.edge_mlp, e, prec)
        vertex_out = P.mlp_rows(self.decoder.vertex_mlp, vx, prec)
        return x, e, vx, edge_out, vertex_out

    def forward(self, graphs, mode="rollout"):   # VertPot.py:74-101
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        _, _, _, edge_attr_out, vertex_out = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        cell_flux = cell_flux_from_vertices(vertex_out, v_graph.face)
        self.dt = c_graph.dt
        acc_pred = self.integrator([cell_flux, edge_attr_out], c_graph, f_graph, self.dt)
        output = [torch.cat([acc_pred, cell_flux], dim=1), edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "cell_flux": output[0][:, 2:5],
                "face_velocity": output[1][:, 0:2], "face_pressure": output[1][:, 2:3]}

    class Integrator(nn.Module):   # VertPot.py:103-150
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.vol_dt_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
            cell_flux, edge_output = output
            uv, p_face, flux_d = edge_output[:, 0:2], edge_output[:, 2:3], edge_output[:, 3:5]
            coeff = normalize_vol_dt(c_graph.volume, c_graph.edge_index, dt, self.vol_dt_norm)
            phi_a = sum(uv[cf[j]] * cell_flux[:, j:j + 1] * coeff[cf[j]] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d

    class GN_Block(FvgnA.GN_Block):   # VertPot.py:187-210
        family = "vertpot"

        def __init__(self, config, hidden_size):
            super().__init__(config, hidden_size)   # keeps the reference's unused face_block/cell_block params
            self.edge_block = self.Face_Block(config, hidden_size)
            self.node_block = self.Cell_Block(config, hidden_size)
            self.vertex_block = nn.Module()        # parameter-free (VertPot.py:212-222)

    class Decoder(nn.Module):   # VertPot.py:224-231
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.edge_mlp = build_mlp(config, hidden_size, hidden_size, output_sizes[1], norm_layer=False)
            self.vertex_mlp = build_mlp(config, hidden_size, hidden_size, output_sizes[2], norm_layer=False)
