"""VertPot family on the B200 kernels - drop-in for reference ``src/models/VertPot.py`` (VertPotA, B, C, E, G;
D and F call a function the reference does not define, ``fvm.convert_cell_flux_to_face_flux_alt``, and cannot run there):
FvgnA blocks (under ``node_block`` / ``edge_block``) plus a Vertex_Block that sums the face block's raw
output onto vertices; edge + vertex decoder heads (VertPot.py:187-231).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import processor as P
from ..topology import get_topology
from .base import build_mlp, col, n_class_types
from .Flux import FluxA, FluxC, cell_to_face, normalize_vol_dt
from .Fvgn import FvgnA, calc_gradient_tensor, flux_dot, graph_topology, normalize_face_area

import os
USE_GATHER3 = os.environ.get("GNNFD_GATHER3", "1") != "0"      # A/B knob


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

    def loss(self, output, graphs):   # VertPot.py:152-186 (continuity on the cell fluxes; no interior mask on faces)
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        cf = output["cell_flux"]
        div = (cf[:, 0] + cf[:, 1] + cf[:, 2]).unsqueeze(-1)                      # fvm.py:13-19
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fvl = lf(output["face_velocity"], f_graph.y[:, 0:2], None, f_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc + w["face_velocity"] * fvl
                 + w["face_pressure"] * fpl)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}

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
        x, fast = P.encode_cells(self.encoder.cell_mlp, c_x, prec, self.family, self.processer_list, topo.n_cells)
        x, e, vx = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook, fast=fast)
        edge_out = P.mlp_rows(self.decoder.edge_mlp, e, prec)
        vertex_out = P.mlp_rows(self.decoder.vertex_mlp, vx, prec)
        return x, e, vx, edge_out, vertex_out

    def forward(self, graphs, mode="rollout"):   # VertPot.py:74-101
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        c_graph.topology = topo          # the integrator reuses its int32 index tensors
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
            topo = graph_topology(c_graph)
            if topo is not None and edge_output.is_cuda and edge_output.shape[1] == 5 and USE_GATHER3:
                # the fifteen x[cf[j]] gathers of the expression below as two gather3 launches (rows of edge_output; the
                # coefficient / area pair), whose autograd is a fixed-degree sort-free transpose instead of index_put; the
                # arithmetic is the same tensor expression in the same order
                from ..fvm_ops import cell_faces, gather3
                cfs = cell_faces(topo, cf)
                coeff = normalize_vol_dt(c_graph.volume, c_graph.edge_index, dt, self.vol_dt_norm, topo=topo)
                area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm, topo=topo)
                self.face_area = area
                eo = gather3(edge_output, cfs, topo.row, topo.col)                              # [3, N, (u, v, p, d0, d1)]
                ka = gather3(torch.cat([coeff, area], dim=1), cfs, topo.row, topo.col)          # [3, N, (coeff, area)]
                phi_a = sum(eo[j][:, 0:2] * cell_flux[:, j:j + 1] * ka[j][:, 0:1] for j in range(3))
                phi_d = eo[0][:, 3:5] + eo[1][:, 3:5] + eo[2][:, 3:5]
                phi_p = sum(eo[j][:, 2:3] * unv[:, j, :] * ka[j][:, 1:2] for j in range(3))
                return 1.0 * (-phi_a - phi_p / self.rho) + phi_d
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


def cell_flux_to_owner_face_flux(cell_flux, edge_index, face_index):
    """Face flux = the owner cell's local flux of that face (utils/fvm.py:55-93; the reference additionally raises when
    a face is not listed exactly once by its owner cell, which a valid mesh never triggers)."""
    owner = edge_index[0]
    mask = face_index[:, owner] == torch.arange(edge_index.shape[1], device=cell_flux.device).unsqueeze(0)
    local = torch.argmax(mask.int(), dim=0)
    return cell_flux[owner, local].unsqueeze(-1)


def cell_flux_to_face_flux_last(cell_flux, face_to_cells, cell_faces):
    """utils/geometry.py:539-570 as written there, including its pairing of ``cell_faces.flatten()`` ([3, N] row-major)
    with cell-major (cell, local face) indices; duplicate targets resolve to the LAST writer, which is what the
    reference's indexed assignment does on the CPU (made explicit here so the GPU result is deterministic)."""
    n, n_faces = cell_flux.shape[0], face_to_cells.shape[1]
    gfi = cell_faces.flatten()
    cells = torch.arange(n, device=cell_flux.device).repeat_interleave(3)
    local = torch.arange(3, device=cell_flux.device).repeat(n)
    flux = cell_flux[cells, local]
    corrected = torch.where(face_to_cells[0, gfi] == cells, flux, -flux)
    pos = torch.full((n_faces,), -1, dtype=torch.long, device=cell_flux.device)
    pos = pos.scatter_reduce(0, gfi, torch.arange(3 * n, device=cell_flux.device), reduce="amax")
    out = torch.where(pos >= 0, corrected[pos.clamp_min(0)], torch.zeros_like(corrected[:1]).expand(n_faces))
    return out.unsqueeze(-1)


class _VertPotNet:
    """Mixin: VertPotA's processor and two-head decoder under another model's glue (VertPot.py:327-334 etc.)."""
    family = "vertpot"
    training_plan = VertPotA.training_plan
    encode_process_decode = VertPotA.encode_process_decode

    def _install_vertpot_net(self, config):
        self.processer_list = nn.ModuleList(
            [VertPotA.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])
        self.decoder = VertPotA.Decoder(config, self.hidden_size, self.output_sizes)

    def _heads(self, graphs):
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, _, edge_out, vertex_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        return edge_out, cell_flux_from_vertices(vertex_out, v_graph.face)


class VertPotB(VertPotA):
    """VertPotA with a physical integrator on de-normalised outputs (VertPot.py:234-319)."""
    face_grad_weights_use = True

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1, nu=1e-3)
        self.face_mls_weights = None

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 3, 1])

    def forward(self, graphs, mode="rollout"):   # VertPot.py:248-281
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, _, edge_attr_out, vertex_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        cell_flux = cell_flux_from_vertices(vertex_out, v_graph.face)
        norm_cell_out = torch.cat([torch.zeros_like(c_graph.x[:, 0:2]), cell_flux], dim=1)
        output = self.normalizer.output([norm_cell_out.clone(), edge_attr_out.clone(), None], inverse=True)
        self.dt = c_graph.dt
        acc_pred = self.integrator(output, c_graph, f_graph, self.dt)
        output[0] = torch.cat([acc_pred, output[0][:, 2:]], dim=1)
        if mode != "rollout":
            output = self.normalizer.output([torch.cat([acc_pred, torch.zeros_like(cell_flux)], dim=1), None, None])
            output[1] = edge_attr_out
            output[0][:, 2:5] = cell_flux
        else:
            output[0][:, 0:2] = acc_pred
        return {"cell_velocity_change": output[0][:, 0:2], "cell_flux": output[0][:, 2:5],
                "face_velocity": output[1][:, 0:2], "face_pressure": output[1][:, 2:3]}

    class Integrator(nn.Module):   # VertPot.py:283-319
        def __init__(self, config, rho, nu=1e-3):
            super().__init__()
            self.rho, self.nu = rho, nu

        def forward(self, output, c_graph, f_graph, dt):
            unv, cf, area = c_graph.normal, f_graph.face, f_graph.area
            uv, p_face, cell_flux = output[1][:, 0:2], output[1][:, 2:3], output[0][:, 2:5]
            phi_a = sum(uv[cf[j]] * cell_flux[:, j:j + 1] for j in range(3))
            grad = calc_gradient_tensor(uv, f_graph.grad_weights, f_graph.grad_neighbours)
            phi_d = sum(flux_dot(grad[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p / self.rho + self.nu * phi_d)


class VertPotC(_VertPotNet, FluxC):
    """Pressure and diffusion on faces, the flux from vertex-potential differences, explicit face velocity
    (VertPot.py:322-444)."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self._install_vertpot_net(config)
        self.integrator = self.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 3, 1])

    def forward(self, graphs, mode="rollout"):   # VertPot.py:340-366
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        edge_attr_out, cell_flux = self._heads(graphs)
        self.dt = c_graph.dt
        acc_pred = self.integrator([cell_flux, edge_attr_out], c_graph, f_graph, self.dt)
        output = [torch.cat([acc_pred, cell_flux], dim=1), edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "cell_flux": output[0][:, 2:5],
                "face_pressure": output[1][:, 0:1]}

    def loss(self, output, graphs):   # VertPot.py:410-444
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        div = output["cell_flux"][:, 0] + output["cell_flux"][:, 1] + output["cell_flux"][:, 2]
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 0:1], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = w["continuity"] * continuity + w["cell_velocity_change"] * cvc + w["face_pressure"] * fpl
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_pressure_loss": fpl}

    class Integrator(nn.Module):   # VertPot.py:368-408
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
            cell_flux, edge_output = output
            uv = cell_to_face(c_graph.x[:, 0:2], c_graph.edge_index, f_graph.pos, c_graph.pos)
            p_face, flux_d = edge_output[:, 0:1], edge_output[:, 1:3]
            phi_a = sum(uv[cf[j]] * cell_flux[:, j:j + 1] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d


class VertPotE(_VertPotNet, FluxC):
    """FluxC's glue and integrator over VertPotA's network; the face flux is read off the owner cell's vertex-potential
    differences and appended to the edge head (VertPot.py:494-539)."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self._install_vertpot_net(config)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 3, 1])

    def forward(self, graphs, mode="rollout"):   # VertPot.py:510-539
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        edge_attr_out, cell_flux = self._heads(graphs)
        face_flux = cell_flux_to_owner_face_flux(cell_flux, c_graph.edge_index, f_graph.face)
        edge_attr_out = torch.cat([edge_attr_out, face_flux], dim=1)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3], "face_flux": output[1][:, 3:4]}


class VertPotG(VertPotA):
    """VertPotA plus a face-flux output (and a loss on it) derived from the per-cell vertex-potential differences
    (VertPot.py:631-818; its GN_Block / Decoder / Integrator restate VertPotA's)."""

    def forward(self, graphs, mode="rollout"):   # VertPot.py:658-687
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, _, edge_attr_out, vertex_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        cell_flux = cell_flux_from_vertices(vertex_out, v_graph.face)
        self.dt = c_graph.dt
        acc_pred = self.integrator([cell_flux, edge_attr_out], c_graph, f_graph, self.dt)
        output = [torch.cat([acc_pred, cell_flux], dim=1), edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        face_flux = cell_flux_to_face_flux_last(output[0][:, 2:5], c_graph.edge_index, f_graph.face)
        return {"cell_velocity_change": output[0][:, 0:2], "face_flux": face_flux,
                "face_velocity": output[1][:, 0:2], "face_pressure": output[1][:, 2:3]}

    def loss(self, output, graphs):   # VertPot.py:737-772
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        ff, flux = f_graph.face, output["face_flux"]
        div = flux[ff[0]] + flux[ff[1]] + flux[ff[2]]
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fvl = lf(output["face_velocity"], f_graph.y[:, 0:2], None, f_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
        ffl = lf(flux, f_graph.y[:, 3:4], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc + w["face_velocity"] * fvl
                 + w["face_pressure"] * fpl + w["face_flux"] * ffl)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}


class _Unrunnable:
    """VertPotD / VertPotF construct and load checkpoints like the reference's classes, but their ``forward`` cannot be
    reproduced: in the reference it raises before producing anything (see the class docstrings)."""
    _why = ""

    def forward(self, graphs, mode="rollout"):
        raise NotImplementedError(f"{type(self).__name__}.forward: {self._why}")


class VertPotD(_Unrunnable, _VertPotNet, FluxA):
    """Reference ``VertPotD`` (VertPot.py:447-492).  Its forward calls ``fvm.convert_cell_flux_to_face_flux_alt``
    (VertPot.py:480), a function ``src/utils/fvm.py`` does not define, so the reference raises AttributeError."""
    _why = ("the reference calls fvm.convert_cell_flux_to_face_flux_alt (VertPot.py:480), which src/utils/fvm.py does "
            "not define; there is no reference behaviour to match")

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self._install_vertpot_net(config)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 5, 1])


class VertPotF(_Unrunnable, _VertPotNet, FluxA):
    """Reference ``VertPotF`` (VertPot.py:541-628).  Same missing function (VertPot.py:574), and its integrator is built
    with ``nu=None`` and multiplies by it (VertPot.py:553, 628)."""
    _why = ("the reference calls fvm.convert_cell_flux_to_face_flux_alt (VertPot.py:574), which src/utils/fvm.py does "
            "not define, and multiplies by nu=None (VertPot.py:553, 628); there is no reference behaviour to match")

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self._install_vertpot_net(config)
        self.integrator = self.Integrator(config, rho=1.0)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 3, 1])

    class Integrator(nn.Module):   # VertPot.py:593-598 (parameter-free)
        def __init__(self, config, rho, nu=None):
            super().__init__()
            self.rho, self.nu = rho, nu
