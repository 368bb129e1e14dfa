"""MGN (FVGN-Direct) family on the B200 kernels - drop-in for reference ``src/models/Mgn.py`` (MgnA).

Block order Face_Block -> Cell_Block (Mgn.py:216-226); node decoder with 3 outputs (Mgn.py:269-275).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import processor as P
from ..graph import Data
from ..topology import get_topology
from .base import Model, build_mlp, col, n_class_types
from .Fvgn import FvgnA


class MgnA(Model):
    family = "mgn"
    cell_grad_weights_use = True

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.encoder = self.Encoder(config, self.input_sizes, self.hidden_size)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)
        self.cell_mls_weights = None   # MovingLeastSquaresWeights is offline preprocessing (out of scope)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [3, 0, 0])   # Mgn.py:59-61

    @classmethod
    def normalisation_tables(cls):   # Mgn.py:97-137
        z = "z_score"
        names = ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
                 "cell_pressure", "face_velocity_difference_x", "face_velocity_difference_y",
                 "face_edge_vector_x", "face_edge_vector_y", "face_area"]
        kinds = {k: z for k in names}
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (1, "x", col(0), "face_velocity_difference_x"), (1, "x", col(1), "face_velocity_difference_y"),
                  (1, "x", col(2), "face_edge_vector_x"), (1, "x", col(3), "face_edge_vector_y"),
                  (1, "x", col(4), "face_area"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (0, "y", col(2), "cell_pressure"),
                  (1, "y", col(0), "cell_velocity_x"), (1, "y", col(1), "cell_velocity_y")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (0, col(2), "cell_pressure")]
        return kinds, inputs, outputs

    def encode_process_decode(self, c_x, f_x, topo, hook=None):
        prec = self.prec
        if self.wants_grad() and self.family in ("fvgn", "mgn"):   # hand-scheduled backward (training.py); other
            # families train through the per-op autograd wrappers the processor is written with (autograd_ops.py)
            from ..training import encode_process_decode_train
            return None, None, encode_process_decode_train(self.training_plan(), topo, prec, c_x, f_x)
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x, fast = P.encode_cells(self.encoder.cell_mlp, c_x, prec, self.family, self.processer_list, topo.n_cells)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook, fast=fast)
        return x, e, P.mlp_rows(self.decoder.face_mlp, x, prec)

    def forward(self, graphs, mode="train"):   # Mgn.py:153-173
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        _, _, cell_output = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        output = [cell_output, None, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3]}

    def update_features(self, output, input_graphs):   # Mgn.py:139-151
        c_graph, f_graph, v_graph = input_graphs
        c_graph.x = output["cell_velocity"].detach()
        u = c_graph.x[:, :2]
        dv = u[c_graph.edge_index[0]] - u[c_graph.edge_index[1]]
        dv = torch.where(f_graph.boundary_mask.unsqueeze(-1), f_graph.y[:, 0:2], dv)   # == dv[mask] = y[mask]
        f_graph.x[:, 0:2] = dv
        return [c_graph, f_graph, v_graph]

    def loss(self, output, graphs):   # Mgn.py:175-197
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        cvc = lf(output["cell_velocity_change"], c_graph.y[:, 0:2], None, c_graph.batch)
        cp = lf(output["cell_pressure"], c_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = w["cell_velocity_change"] * cvc + w["cell_pressure"] * cp
        return {"total_log_loss": torch.mean(torch.log(total)), "cell_velocity_change_loss": cvc,
                "cell_pressure_loss": cp}

    Encoder = FvgnA.Encoder   # identical container (Mgn.py:199-208)

    class GN_Block(FvgnA.GN_Block):   # same sub-modules, Face_Block -> Cell_Block order
        family = "mgn"

    class Decoder(nn.Module):   # Mgn.py:269-275
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.face_mlp = build_mlp(config, hidden_size, hidden_size, output_sizes[0], norm_layer=False)

        def forward(self, graph, prec=0):
            return P.mlp_rows(self.face_mlp, graph.x, prec)


def divergence_from_uc(cell_velocity, weights, neighbours, cell_volume):   # utils/fvm.py:40-52
    ux, uy = cell_velocity[:, 0], cell_velocity[:, 1]
    gx = torch.sum(weights[:, :, 0] * (ux[neighbours] - ux[:, None]), dim=1)
    gy = torch.sum(weights[:, :, 1] * (uy[neighbours] - uy[:, None]), dim=1)
    return (gx + gy).unsqueeze(-1) * cell_volume


class MgnB(MgnA):
    """Direct prediction of the next velocity, z-score normalisation (Mgn.py:278-391): same encoder / processor /
    decoder kernels as MgnA, different normalisation tables, output keys and loss."""

    @classmethod
    def normalisation_tables(cls):   # Mgn.py:318-337
        kinds, inputs, outputs = MgnA.normalisation_tables()
        inputs = [r for r in inputs if r[3] not in ("cell_velocity_change_x", "cell_velocity_change_y")]
        inputs += [(0, "y", col(0), "cell_velocity_x"), (0, "y", col(1), "cell_velocity_y")]
        outputs = [(0, col(0), "cell_velocity_x"), (0, col(1), "cell_velocity_y"), (0, col(2), "cell_pressure")]
        return kinds, inputs, outputs

    def forward(self, graphs, mode="train"):   # Mgn.py:340-360
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        _, _, cell_output = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        output = [cell_output, None, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3]}

    def loss(self, output, graphs):   # Mgn.py:362-391
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        div = divergence_from_uc(output["cell_velocity"], c_graph.grad_weights, c_graph.grad_neighbours, c_graph.volume)
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cv = lf(output["cell_velocity"], c_graph.y[:, 0:2], None, c_graph.batch)
        cp = lf(output["cell_pressure"], c_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = w["cell_velocity"] * cv + w["cell_pressure"] * cp + w["continuity"] * continuity
        return {"total_log_loss": torch.mean(torch.log(total)), "cell_velocity_loss": cv,
                "cell_pressure_loss": cp, "continuity_loss": continuity}


class MgnC(MgnB):
    """MgnB with the velocities scaled by the characteristic velocity (mean_scale) instead of z-scored (Mgn.py:394-424)."""

    @classmethod
    def normalisation_tables(cls):   # Mgn.py:402-424
        kinds, inputs, outputs = MgnB.normalisation_tables()
        kinds["cell_velocity_char"] = "mean_scale"
        swap = lambda r: r[:-1] + ("cell_velocity_char",) if r[0] == 0 and r[-1] in ("cell_velocity_x", "cell_velocity_y") else r
        inputs = [swap(r) for r in inputs]          # graphs[0].x / .y velocity columns; the face B.C. rows keep z-score
        outputs = [(0, col(0), "cell_velocity_char"), (0, col(1), "cell_velocity_char"), (0, col(2), "cell_pressure")]
        return kinds, inputs, outputs
