"""Stream-function (divergence-free by construction) variants of MGN on the B200 kernels - drop-in for reference
``src/models/StreamFunc.py``.  The encoder / 15 GN_Blocks / decoder are MgnA's (family "mgn"); the decoder predicts a
scalar potential + pressure, and the velocity is the rotated moving-least-squares gradient of the potential.
"""
from __future__ import annotations

import torch
from torch import nn

from ..topology import get_topology
from .Mgn import MgnB, MgnC, divergence_from_uc
from .base import n_class_types

INFLOW, WALL_BOUNDARY = 2, 1   # datasets/OpenFoam.py:19-24


class DivergenceLayer(nn.Module):   # StreamFunc.py:94-106
    def forward(self, cell_potential, weights, neighbours):
        diff = cell_potential[neighbours] - cell_potential[:, None]
        gx = torch.sum(weights[:, :, 0] * diff, dim=1)
        gy = torch.sum(weights[:, :, 1] * diff, dim=1)
        return torch.stack([-gy, gx], dim=1)


class SmoothingLayer(nn.Module):   # StreamFunc.py:277-287
    def __init__(self, neighbours=3):
        super().__init__()
        self.neighbours = neighbours

    def forward(self, potential, neighbours):
        return torch.mean(potential[neighbours[:, :self.neighbours]], dim=1)


class BaseStreamFunc:   # StreamFunc.py:32-91
    DivergenceLayer = DivergenceLayer

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.divergence_layer = DivergenceLayer()
        self.cell_grad_weights_use = True
        self.cell_mls_weights = None   # MovingLeastSquaresWeights is offline preprocessing (out of scope)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [2, 0, 0])

    def _decode(self, graphs):
        """encoder -> 15 GN_Blocks -> decoder on the (already normalised) graphs: [N, 2] = (potential, pressure)."""
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, cell_output = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        return cell_output

    def loss(self, output, graphs):   # StreamFunc.py:45-75
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        div = divergence_from_uc(output["cell_velocity"], c_graph.grad_weights, c_graph.grad_neighbours, c_graph.volume)
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cv = lf(output["cell_velocity"], c_graph.y[:, 0:2], None, c_graph.batch)
        cp = lf(output["cell_pressure"], c_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = w["cell_velocity"] * cv + w["cell_pressure"] * cp
        return {"total_log_loss": torch.mean(torch.log(total)), "cell_velocity_loss": cv,
                "cell_pressure_loss": cp, "continuity_loss": continuity}

    def update_features(self, output, input_graphs):   # StreamFunc.py:77-91
        c_graph, f_graph, v_graph = input_graphs
        c_graph.x = output["cell_velocity"].detach()
        u = c_graph.x[:, :2]
        dv = u[c_graph.edge_index[0]] - u[c_graph.edge_index[1]]
        mask = ((f_graph.type == INFLOW) | (f_graph.type == WALL_BOUNDARY)).reshape(-1, 1)
        f_graph.x[:, 0:2] = torch.where(mask, f_graph.y[:, 0:2], dv)
        return [c_graph, f_graph, v_graph]


class StreamFuncA(BaseStreamFunc, MgnC):
    """Velocity = rotated gradient of the potential in NORMALISED space (StreamFunc.py:109-135)."""

    def forward(self, graphs, mode="rollout"):
        graphs = self.normalizer.input(graphs)
        c_graph = graphs[0]
        cell_output = self._decode(graphs)
        u = self.divergence_layer(cell_output[:, 0], c_graph.grad_weights, c_graph.grad_neighbours)
        output = [torch.cat([u, cell_output[:, 1:2]], dim=1), None, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3]}


class StreamFuncB(BaseStreamFunc, MgnC):
    """Velocity from the DE-normalised potential, re-normalised for the training loss (StreamFunc.py:138-167)."""

    def _potential(self, cell_output, grad_neighbours):
        return cell_output[:, 0:1]

    def forward(self, graphs, mode="rollout"):
        graphs = self.normalizer.input(graphs)
        c_graph = graphs[0]
        cell_output = self._decode(graphs)
        raw_potential = cell_output[:, 0:1].clone()
        potential = self._potential(cell_output, c_graph.grad_neighbours)
        expanded = torch.cat([potential, torch.zeros_like(cell_output[:, 0:1]), cell_output[:, 1:2]], dim=1)
        output = self.normalizer.output([expanded, None, None], inverse=True)
        u = self.divergence_layer(output[0][:, 0], c_graph.grad_weights, c_graph.grad_neighbours)
        output[0][:, 0:2] = u
        if mode == "train":
            output = self.normalizer.output(output, inverse=False)
        return self._result(output, raw_potential)

    def _result(self, output, raw_potential):
        return {"cell_velocity": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3]}


class StreamFuncC(BaseStreamFunc, MgnB):
    """No normalisation inside forward (StreamFunc.py:170-192)."""

    def forward(self, graphs, mode="rollout"):
        c_graph = graphs[0]
        cell_output = self._decode(graphs)
        u = self.divergence_layer(cell_output[:, 0], c_graph.grad_weights, c_graph.grad_neighbours)
        return {"cell_velocity": u, "cell_pressure": cell_output[:, 1:2]}


class StreamFuncD(StreamFuncB):
    """StreamFuncB with the potential averaged over 8 neighbours before differentiation; the raw potential is returned
    for a smoothness regulariser (StreamFunc.py:195-275)."""
    SmoothingLayer = SmoothingLayer

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.smoother = SmoothingLayer(neighbours=8)

    def _potential(self, cell_output, grad_neighbours):
        return self.smoother(cell_output[:, 0], grad_neighbours)[:, None]

    def _result(self, output, raw_potential):
        return {"cell_velocity": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3],
                "cell_potential": raw_potential}

    def loss(self, output, graphs):   # StreamFunc.py:237-275
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        div = divergence_from_uc(output["cell_velocity"], c_graph.grad_weights, c_graph.grad_neighbours, c_graph.volume)
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cv = lf(output["cell_velocity"], c_graph.y[:, 0:2], None, c_graph.batch)
        cp = lf(output["cell_pressure"], c_graph.y[:, 2:3], None, f_graph.batch)
        potential = output["cell_potential"]
        lap = torch.mean(potential[c_graph.grad_neighbours[:, :4]], dim=1) - potential
        smooth = torch.mean(lap ** 2)
        w = self.config.training.loss_weights
        total = w["cell_velocity"] * cv + w["cell_pressure"] * cp + 0.1 * smooth
        return {"total_log_loss": torch.mean(torch.log(total)), "cell_velocity_loss": cv,
                "cell_pressure_loss": cp, "continuity_loss": continuity}
