"""FVGN family on the B200 kernels - drop-in for reference ``src/models/Fvgn.py`` (FvgnA).

Same public surface (constructor, ``forward(graphs, mode)`` dict keys, ``loss``, ``update_features``,
state_dict keys ``encoder.{face_mlp,cell_mlp}``, ``processer_list.{i}.{face_block.face_mlp,
cell_block.cell_mlp}``, ``decoder.face_mlp``, ``integrator.face_area_norm``).  Encoder, the 15
GN_Blocks and the decoder run through ``processor.py`` (CUDA kernels); the finite-volume integrator,
loss and rollout glue are plain tensor code ("next" rows).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import processor as P
from ..graph import Data
from ..mesh import NODE_INFLOW, NODE_WALL
from ..topology import get_topology
from .base import Model, build_mlp, col, n_class_types


def normalize_face_area(face_area, cell_volume, edge_index, dt, batch_norm, topo=None):
    """face_area * mean(dt) / mean adjacent cell volume, through BatchNorm1d(1)
    (reference utils/normalisation.py:325-344).  With the mesh topology at hand (int32 row / col on the GPU) this is
    the fused kernel pair of ``fvm_ops.face_area_norm`` (batch statistics, running-stat update and normalisation)."""
    if topo is not None and face_area.is_cuda and isinstance(batch_norm, nn.BatchNorm1d):
        from ..fvm_ops import face_area_norm
        return face_area_norm(face_area, cell_volume, topo.row, topo.col, dt, batch_norm)
    vol = (cell_volume.index_select(0, edge_index[0]) + cell_volume.index_select(0, edge_index[1])) / 2
    return batch_norm((face_area * (torch.mean(dt) / vol)).view(-1, 1))


def graph_topology(c_graph):
    """The MeshTopology ``forward`` attached to the batch (None for a hand-built call without one)."""
    from ..topology import MeshTopology
    topo = getattr(c_graph, "topology", None)
    if isinstance(topo, MeshTopology) and topo.n_faces == c_graph.edge_index.shape[1] and topo.n_cells == c_graph.x.shape[0]:
        return topo
    return None


def flux_dot(a, n):
    """Pairs of columns of ``a`` dotted with the 2-vector ``n`` (reference utils/maths.py:12-20)."""
    return torch.cat([(a[:, i:i + 2] * n).sum(-1, keepdim=True) for i in range(0, a.size(1), 2)], dim=-1)


class FvgnA(Model):
    family = "fvgn"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.encoder = self.Encoder(config, self.input_sizes, self.hidden_size)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)
        self.integrator = self.Integrator(config, rho=1)

    # --- interface classmethods ---------------------------------------------------------------
    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 5, 0])   # Fvgn.py:51-53

    @classmethod
    def normalisation_tables(cls):   # Fvgn.py:55-99
        z = "z_score"
        names = ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
                 "face_velocity_difference_x", "face_velocity_difference_y", "face_edge_vector_x",
                 "face_edge_vector_y", "face_area", "face_velocity_x", "face_velocity_y", "face_pressure"]
        kinds = {k: z for k in names}
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (1, "x", col(0), "face_velocity_difference_x"), (1, "x", col(1), "face_velocity_difference_y"),
                  (1, "x", col(2), "face_edge_vector_x"), (1, "x", col(3), "face_edge_vector_y"),
                  (1, "x", col(4), "face_area"),
                  (1, "y", col(0), "face_velocity_x"), (1, "y", col(1), "face_velocity_y"),
                  (1, "y", col(2), "face_pressure")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (1, col(0), "face_velocity_x"), (1, col(1), "face_velocity_y"),
                   (1, col(2), "face_pressure")]
        return kinds, inputs, outputs

    # --- hot path ---------------------------------------------------------------------------------
    def encode_process_decode(self, c_x, f_x, topo, hook=None):
        """encoder -> mp_num GN_Blocks -> decoder on normalised inputs; returns (x, e, decoder out)."""
        prec = self.prec
        if self.wants_grad() and self.family in ("fvgn", "mgn"):   # hand-scheduled backward (training.py); other
            # families train through the per-op autograd wrappers the processor is written with (autograd_ops.py)
            from ..training import encode_process_decode_train
            return None, None, encode_process_decode_train(self.training_plan(), topo, prec, c_x, f_x)
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x, fast = P.encode_cells(self.encoder.cell_mlp, c_x, prec, self.family, self.processer_list, topo.n_cells)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook, fast=fast)
        return x, e, P.mlp_rows(self.decoder.face_mlp, e, prec)

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:150-174
        return self.forward_normalised(self.normalizer.input(graphs), mode)

    def forward_normalised(self, graphs, mode="rollout"):
        """``forward`` after the in-place input normalisation (graphs already normalised and resident)."""
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
        c_graph.topology = topo          # the integrator and the loss of this batch reuse its int32 index tensors
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, topo)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2],
                "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    # --- glue (plain tensor code) -----------------------------------------------------------------
    def update_features(self, output, input_graphs):   # Fvgn.py:133-148
        c_graph, f_graph, v_graph = input_graphs
        c_graph.x = output["cell_velocity"].detach()
        u = c_graph.x[:, :2]
        dv = u[c_graph.edge_index[0]] - u[c_graph.edge_index[1]]
        mask = ((f_graph.type == NODE_INFLOW) | (f_graph.type == NODE_WALL)).squeeze(-1)
        dv = torch.where(mask.unsqueeze(-1), f_graph.y[:, 0:2], dv)   # == dv[mask] = y[mask], without the host sync
        f_graph.x[:, 0:2] = dv
        return [c_graph, f_graph, v_graph]

    def loss(self, output, graphs):   # Fvgn.py:176-212
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        topo = graph_topology(c_graph)
        face_area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, self.dt,
                                        self.integrator.face_area_norm, topo=topo)
        ff, unv, fv = f_graph.face, c_graph.normal, output["face_velocity"]
        if topo is not None and fv.is_cuda:      # fused fixed-degree divergence (fvm.py:26-37) with its own backward
            from ..fvm_ops import cell_faces, fvm_divergence
            div = fvm_divergence(fv, face_area, unv, cell_faces(topo, ff), topo.row, topo.col)
        else:
            div = sum(flux_dot(fv[ff[j]], unv[:, j, :]) * face_area[ff[j]] for j in range(3))
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fvl = lf(output["face_velocity"], f_graph.y[:, :2], ~f_graph.boundary_mask, f_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc
                 + w["face_velocity"] * fvl + w["face_pressure"] * fpl)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}

    # --- parameter containers (reference layout) --------------------------------------------------
    class Integrator(nn.Module):   # Fvgn.py:214-255
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf = c_graph.normal, f_graph.face
            topo = graph_topology(c_graph)
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm, topo=topo)
            self.face_area = area
            if topo is not None and edge_output.is_cuda and edge_output.shape[1] == 5:
                # one kernel forward, one backward (fixed-degree gathers) instead of ~20 tensor kernels
                from ..fvm_ops import cell_faces, fvm_integrate
                return fvm_integrate(edge_output, area, unv, cell_faces(topo, cf), topo.row, topo.col, self.rho)
            uv, p_face, flux_d = edge_output[:, :2], edge_output[:, 2:3], edge_output[:, 3:]
            uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)
            phi_a = sum(flux_dot(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
            phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d

    class Encoder(nn.Module):   # Fvgn.py:257-266
        def __init__(self, config, input_sizes, hidden_size):
            super().__init__()
            self.face_mlp = build_mlp(config, input_sizes[1], hidden_size, hidden_size)
            self.cell_mlp = build_mlp(config, input_sizes[0], hidden_size, hidden_size)

        def forward(self, cell_graph, prec=0):
            return Data(x=P.mlp_rows(self.cell_mlp, cell_graph.x, prec),
                        edge_attr=P.mlp_rows(self.face_mlp, cell_graph.edge_attr, prec),
                        edge_index=cell_graph.edge_index)

    class GN_Block(nn.Module):   # Fvgn.py:268-325
        family = "fvgn"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block = self.Face_Block(config, hidden_size)
            self.cell_block = self.Cell_Block(config, hidden_size)

        def forward(self, c_graph, v_graph, prec=0):
            """Reference call shape ``gnblock(c_graph, v_graph) -> Data(x, edge_attr, edge_index)``."""
            topo = get_topology([c_graph, None, v_graph])
            x, e, _ = P.gn_block(self.family, self, c_graph.x, c_graph.edge_attr, topo, prec)
            return Data(x=x, edge_attr=e, edge_index=c_graph.edge_index)

        class Face_Block(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp(config, hidden_size * 3, hidden_size, hidden_size)

        class Cell_Block(nn.Module):
            def __init__(self, config, hidden_size, mp_times=2):
                super().__init__()
                self.cell_mlp = build_mlp(config, hidden_size + hidden_size // 2, hidden_size, hidden_size)
                self.mp_times = mp_times

    class Decoder(nn.Module):   # Fvgn.py:327-333
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.face_mlp = build_mlp(config, hidden_size, hidden_size, output_sizes[1], norm_layer=False)

        def forward(self, graph, prec=0):
            return P.mlp_rows(self.face_mlp, graph.edge_attr, prec)


class FvgnF(FvgnA):
    """Reference ``FvgnF`` (Fvgn.py:881-1002): ONE GN_Block whose weights are shared by all ``mp_num`` message-passing
    steps; every MLP input carries an extra constant column ``(step + 1) / mp_num``.  A constant input column is a
    bias: ``W1 [in | c] + b1 = W1[:, :-1] in + (b1 + c W1[:, -1])``, so the kernels run the ordinary K = 384 / 192
    blocks with a per-step effective first-layer bias and the shared operand pack (state_dict layout unchanged:
    ``gn_block.{face,cell}_block.*`` with 385 / 193 input columns, plus FvgnA's unused ``processer_list``)."""
    family = "fvgn_f"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.encoder = self.Encoder(config, self.input_sizes, self.hidden_size)
        self.gn_block = self.GN_Block(config, self.hidden_size)
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)
        self.mp_num = config.model.mp_num

    def _step_weights(self, seq):
        """Per-step MLPWeights of a shared MLP: W1 without its last column, b1 + step * W1[:, -1]; cached until a
        parameter changes."""
        inner, ln = P._split_mlp(seq)
        lin = [m for m in inner if isinstance(m, nn.Linear)]
        params = [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias, ln.weight, ln.bias]
        key = tuple((p.data_ptr(), p._version) for p in params)
        cached = getattr(seq, "_gnnfd_step_w", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        w1 = lin[0].weight.detach()
        w1_main, w1_step = w1[:, :-1].contiguous(), w1[:, -1]
        out = []
        for i in range(self.mp_num):
            b1 = (lin[0].bias.detach() + ((i + 1) / self.mp_num) * w1_step).contiguous()
            out.append(P.MLPWeights(w1=w1_main, b1=b1, w2=lin[1].weight.detach(), b2=lin[1].bias.detach(),
                                    w3=lin[2].weight.detach(), b3=lin[2].bias.detach(), ln_w=ln.weight.detach(),
                                    ln_b=ln.bias.detach(), has_ln=True, ln_eps=ln.eps))
        object.__setattr__(seq, "_gnnfd_step_w", (key, out))
        return out

    def encode_process_decode(self, c_x, f_x, topo, hook=None):
        prec = self.prec
        if self.wants_grad():
            raise NotImplementedError("FvgnF runs forward / rollout only on the B200 path (wrap the call in torch.no_grad())")
        ops, Seg = P.ops, P.Seg
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        wn_steps = self._step_weights(self.gn_block.cell_block.cell_mlp)
        we_steps = self._step_weights(self.gn_block.face_block.face_mlp)
        for i in range(self.mp_num):
            wn, we = wn_steps[i], we_steps[i]
            for w, first in ((wn, wn_steps[0]), (we, we_steps[0])):
                if w is not first and first.packed is not None and w.packed is None:
                    w.packed, w.packed_prec = first.packed, first.packed_prec
            vsum = P.vertex_half_sum(e, topo)
            x_raw, x_new = ops.mlp_forward([Seg(x), Seg(vsum, P.SEG_MEAN3, topo.vf)], wn, x.shape[0], prec,
                                           residual=x, want_raw=True, want_sum=True)
            _, e_new = ops.mlp_forward([Seg(e), Seg(x_raw, P.SEG_GATHER, (topo.row,)), Seg(x_raw, P.SEG_GATHER, (topo.col,))],
                                       we, e.shape[0], prec, residual=e, want_raw=False, want_sum=True)
            x, e = x_new, e_new
            if hook is not None:
                hook(i, x, e)
        return x, e, P.mlp_rows(self.decoder.face_mlp, e, prec)

    class GN_Block(nn.Module):   # Fvgn.py:940-996
        family = "fvgn_f"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block = self.Face_Block(config, hidden_size)
            self.cell_block = self.Cell_Block(config, hidden_size)

        class Face_Block(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp(config, hidden_size * 3 + 1, hidden_size, hidden_size)

        class Cell_Block(nn.Module):
            def __init__(self, config, hidden_size, mp_times=2):
                super().__init__()
                self.cell_mlp = build_mlp(config, hidden_size + hidden_size // 2 + 1, hidden_size, hidden_size)
                self.mp_times = mp_times


# ------------------------------------------------------------------------------------------------------------------
# Glue-only variants: FvgnA's encoder / 15 GN_Blocks / decoder (the CUDA hot path) with other normalisation tables,
# output scalings, integrators and losses.
# ------------------------------------------------------------------------------------------------------------------
def calc_gradient_tensor(value, weights, neighbours):   # utils/geometry.py:520-537 (index pairing as written there)
    vx, vy = value[:, 0], value[:, 1]
    dx, dy = vx[neighbours] - vx[:, None], vy[neighbours] - vy[:, None]
    return torch.stack([torch.sum(weights[:, :, 0] * dx, dim=1), torch.sum(weights[:, :, 1] * dy, dim=1),
                        torch.sum(weights[:, :, 0] * dy, dim=1), torch.sum(weights[:, :, 1] * dx, dim=1)], dim=1)


def _advection_pressure(edge_output, c_graph, f_graph):
    """Phi_A (u u . n A) and Phi_P (p n A) with the PHYSICAL face area (Fvgn.py:431-460 and 1244-1273)."""
    unv, cf, area = c_graph.normal, f_graph.face, f_graph.area
    uv, p_face = edge_output[:, 0:2], edge_output[:, 2:3]
    uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)
    phi_a = sum(flux_dot(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
    phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
    return phi_a, phi_p


def _real_space_loss(model, output, graphs):
    """Loss of the real-space variants B / J / K (Fvgn.py:388-423 == 1201-1236 == 1343-1378): the continuity term uses
    the NORMALISED face area column of the face features."""
    c_graph, f_graph, v_graph = graphs
    lf = model.mse_term
    ff, unv, fv, area = f_graph.face, c_graph.normal, output["face_velocity"], f_graph.x[:, 4:5]
    div = sum(flux_dot(fv[ff[j]], unv[:, j, :]) * area[ff[j]] for j in range(3))
    continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
    cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
    fvl = lf(output["face_velocity"], f_graph.y[:, :2], ~f_graph.boundary_mask, f_graph.batch)
    fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
    w = model.config.training.loss_weights
    total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc
             + w["face_velocity"] * fvl + w["face_pressure"] * fpl)
    return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
            "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}


class FvgnB(FvgnA):
    """Real-space integration: the decoder output is de-normalised before the integrator, diffusion comes from a
    moving-least-squares velocity gradient instead of a predicted flux (Fvgn.py:336-460).  Decoder width 3."""
    face_grad_weights_use = True

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1, nu=1e-3)
        self.face_mls_weights = None   # MovingLeastSquaresWeights is offline preprocessing (out of scope)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 5 + n_class_types(dataset), 0], [0, 3, 0])

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:360-386
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        output = self.normalizer.output([None, edge_attr_out.clone(), None], inverse=True)
        self.dt = c_graph.dt
        acc_pred = self.integrator(output[1], c_graph, f_graph, self.dt)
        output = [acc_pred, output[1], None]
        if mode == "train":
            output = self.normalizer.output(output)
        return {"cell_velocity_change": output[0], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def loss(self, output, graphs):
        return _real_space_loss(self, output, graphs)

    class Integrator(nn.Module):   # Fvgn.py:425-460
        def __init__(self, config, rho, nu=None):
            super().__init__()
            self.rho, self.nu = rho, nu

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf, area = c_graph.normal, f_graph.face, f_graph.area
            phi_a, phi_p = _advection_pressure(edge_output, c_graph, f_graph)
            grad = calc_gradient_tensor(edge_output[:, :2], f_graph.grad_weights, f_graph.grad_neighbours)
            phi_d = sum(flux_dot(grad[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p / self.rho + self.nu * phi_d)


class FvgnC(FvgnA):
    """Temporal bundling: the decoder predicts ``bundle_size`` future steps at once, [E, k, 5] (Fvgn.py:463-786; its
    Encoder / GN_Block restate FvgnA's).  The normalisation tables are FvgnA's applied on the last dimension."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        window = config.model.bundle_size
        self.output_sizes = [o * window for o in self.output_sizes]
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)
        self.integrator = self.Integrator(config, rho=1)

    def training_plan(self):
        raise NotImplementedError("FvgnC trains through the per-op autograd wrappers (autograd_ops.py)")

    def encode_process_decode(self, c_x, f_x, topo, hook=None):
        prec = self.prec
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x, fast = P.encode_cells(self.encoder.cell_mlp, c_x, prec, self.family, self.processer_list, topo.n_cells)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook, fast=fast)
        n_out = self.output_sizes[1]
        if n_out <= 16:                       # narrow-head path of the fused kernel
            out = P.mlp_rows(self.decoder.face_mlp, e, prec)
        else:                                 # wider bundles: last Linear zero-padded to the 128-wide path (no grad)
            if self.wants_grad():
                raise NotImplementedError("FvgnC with 5 * bundle_size > 16 runs forward / rollout only on the B200 path")
            from .Conservative import _padded_head
            from .._lib import ACT_SILU
            raw, _ = P.ops.mlp_forward([P.Seg(e)], _padded_head(self.decoder.face_mlp, ACT_SILU, n_out), e.shape[0], prec)
            out = raw[:, :n_out].contiguous()
        return x, e, out.view(out.shape[0], n_out // 5, 5)      # Fvgn.py:780-786

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:572-596
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, :, 0:2], "face_velocity": output[1][:, :, 0:2],
                "face_pressure": output[1][:, :, 2:3]}

    def update_features(self, output, input_graphs):   # Fvgn.py:546-560: boundary values of the LAST bundled target
        c_graph, f_graph, v_graph = input_graphs
        c_graph.x = output["cell_velocity"].detach()
        u = c_graph.x[:, :2]
        dv = u[c_graph.edge_index[0]] - u[c_graph.edge_index[1]]
        mask = ((f_graph.type == NODE_INFLOW) | (f_graph.type == NODE_WALL)).reshape(-1, 1)
        f_graph.x[:, 0:2] = torch.where(mask, f_graph.y[:, -1, 0:2], dv)
        return [c_graph, f_graph, v_graph]

    def loss(self, output, graphs):   # Fvgn.py:598-653: per-step losses averaged over the bundle
        c_graph, f_graph, v_graph = graphs
        lf, w = self.mse_term, self.config.training.loss_weights
        ff, unv = f_graph.face, c_graph.normal
        parts = {"total": [], "continuity": [], "cvc": [], "fv": [], "fp": []}
        for t in range(output["face_velocity"].shape[1]):
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, self.dt,
                                       self.integrator.face_area_norm)
            fv = output["face_velocity"][:, t, :]
            div = sum(flux_dot(fv[ff[j]], unv[:, j, :]) * area[ff[j]] for j in range(3))
            continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
            cvc = lf(output["cell_velocity_change"][:, t, :], c_graph.y[:, t, :], None, c_graph.batch)
            fvl = lf(fv, f_graph.y[:, t, :2], ~f_graph.boundary_mask, f_graph.batch)
            fpl = lf(output["face_pressure"][:, t, :], f_graph.y[:, t, 2:3], None, f_graph.batch)
            parts["total"].append(w["continuity"] * continuity + w["cell_velocity_change"] * cvc
                                  + w["face_velocity"] * fvl + w["face_pressure"] * fpl)
            for k, v in (("continuity", continuity), ("cvc", cvc), ("fv", fvl), ("fp", fpl)):
                parts[k].append(v)
        mean = lambda k: torch.mean(torch.stack(parts[k]))
        return {"total_log_loss": torch.mean(torch.log(mean("total"))), "continuity_loss": mean("continuity"),
                "cell_velocity_change_loss": mean("cvc"), "face_velocity_loss": mean("fv"),
                "face_pressure_loss": mean("fp")}

    class Integrator(nn.Module):   # Fvgn.py:655-703: FvgnA's update per bundled step, scaled by (k + 1)
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            unv, cf, k = c_graph.normal, f_graph.face, edge_output.shape[1]
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            results = []
            for t in range(k):
                uv, p_face, flux_d = edge_output[:, t, :2], edge_output[:, t, 2:3], edge_output[:, t, 3:]
                uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)
                phi_a = sum(flux_dot(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
                phi_d = flux_d[cf[0], :] + flux_d[cf[1], :] + flux_d[cf[2], :]
                phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
                results.append((1.0 * (-phi_a - phi_p / self.rho) + phi_d) * (k + 1))
            return torch.stack(results, dim=1)


class FvgnD(FvgnA):
    """Push-forward training: same network and forward as FvgnA; the dataset side changes the target and the
    statistics (Fvgn.py:789-836)."""
    pushforward_use = True

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.pushforward_use = True


class FvgnE(FvgnA):
    """Physical normalisation: every quantity scaled by a characteristic velocity / length / pressure
    (Fvgn.py:839-880)."""

    @classmethod
    def normalisation_tables(cls):
        kinds = {"characteristic_velocity": "max_scale", "characteristic_length": "mean_scale",
                 "characteristic_pressure": "max_scale"}
        u, l, pr = "characteristic_velocity", "characteristic_length", "characteristic_pressure"
        inputs = [(0, "x", col(0), u), (0, "x", col(1), u), (0, "y", col(0), u), (0, "y", col(1), u),
                  (1, "x", col(0), u), (1, "x", col(1), u), (1, "x", col(2), l), (1, "x", col(3), l),
                  (1, "x", col(4), l), (1, "y", col(0), u), (1, "y", col(1), u), (1, "y", col(2), pr)]
        outputs = [(0, col(0), u), (0, col(1), u), (1, col(0), u), (1, col(1), u), (1, col(2), pr)]
        return kinds, inputs, outputs


class FvgnH(FvgnA):
    """Augmented face features: [du(2), face normal(2), area, adjacent distance, angle, one-hot] = 7 + |NodeType|
    columns (Fvgn.py:1013-1114)."""

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 7 + n_class_types(dataset), 0], [0, 5, 0])

    @classmethod
    def normalisation_tables(cls):   # Fvgn.py:1064-1114
        z = "z_score"
        names = ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
                 "face_normal_x", "face_normal_y", "face_area", "face_adjacent_distance", "face_angle",
                 "face_velocity_x", "face_velocity_y", "face_pressure", "face_velocity_difference_x",
                 "face_velocity_difference_y"]
        kinds = {k: z for k in names}
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (1, "x", col(0), "face_velocity_difference_x"), (1, "x", col(1), "face_velocity_difference_y"),
                  (1, "x", col(4), "face_area"), (1, "x", col(5), "face_adjacent_distance"),
                  (1, "x", col(6), "face_angle"), (1, "x", col(2), "face_normal_x"), (1, "x", col(3), "face_normal_y"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (1, "y", col(0), "face_velocity_x"), (1, "y", col(1), "face_velocity_y"),
                  (1, "y", col(2), "face_pressure")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (1, col(0), "face_velocity_x"), (1, col(1), "face_velocity_y"), (1, col(2), "face_pressure")]
        return kinds, inputs, outputs


class FvgnI(FvgnA):
    """FvgnA whose rollout feature update re-imposes INFLOW and WALL faces (Fvgn.py:1117-1137); FvgnA.update_features
    here already implements exactly that masking, so only the class identity differs."""


class FvgnJ(FvgnA):
    """Learnt real-space output scales / biases and a physical integrator (Fvgn.py:1140-1273)."""
    face_grad_weights_use = False

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1, nu=1e-3)
        self.face_mls_weights = None
        self.velocity_scale_x = nn.Parameter(torch.tensor(1.0))
        self.velocity_scale_y = nn.Parameter(torch.tensor(0.01))
        self.pressure_scale = nn.Parameter(torch.tensor(1.0))
        self.diffusion_scale = nn.Parameter(torch.tensor(1.0))
        self.velocity_bias_x = nn.Parameter(torch.tensor(0.0))
        self.velocity_bias_y = nn.Parameter(torch.tensor(0.0))
        self.pressure_bias = nn.Parameter(torch.tensor(0.0))
        self.diffusion_bias = nn.Parameter(torch.tensor(0.0))

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:1164-1199
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, raw = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        edge_attr_out = torch.cat([raw[:, 0:1] * self.velocity_scale_x + self.velocity_bias_x,
                                   raw[:, 1:2] * self.velocity_scale_y + self.velocity_bias_y,
                                   raw[:, 2:3] * self.pressure_scale + self.pressure_bias,
                                   raw[:, 3:5] * self.diffusion_scale + self.diffusion_bias], dim=-1)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode != "rollout":
            output = self.normalizer.output(output)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def loss(self, output, graphs):
        return _real_space_loss(self, output, graphs)

    class Integrator(nn.Module):   # Fvgn.py:1238-1273
        def __init__(self, config, rho, nu=None):
            super().__init__()
            self.rho, self.nu = rho, nu

        def forward(self, edge_output, c_graph, f_graph, dt):
            cf, q = f_graph.face, edge_output[:, 3:5]
            phi_a, phi_p = _advection_pressure(edge_output, c_graph, f_graph)
            phi_d = q[cf[0], :] + q[cf[1], :] + q[cf[2], :]
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p / self.rho + self.nu * phi_d)


class FvgnK(FvgnA):
    """Dimensionless outputs scaled per mesh by the inflow velocity and the Reynolds length (Fvgn.py:1276-1416)."""
    host_sync_in_forward = True     # boolean-mask indexing per mesh: RolloutEngine steps it eagerly (no graph capture)

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1, nu=1e-3)
        self.face_mls_weights = None
        self.anisotropy_ratio = nn.Parameter(torch.tensor(0.0001))

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:1290-1341
        c0, f0 = graphs[0], graphs[1]
        inflow = (f0.type == NODE_INFLOW)
        u_ref = []
        for b in torch.unique(c0.batch):          # reference values are read BEFORE the in-place normalisation
            m = (f0.batch == b).reshape(inflow.shape) & inflow
            if m.any():
                u_ref.append(f0.y[m.reshape(-1)][0, 0])
            else:
                u_ref.append(torch.tensor(1.0, device=c0.y.device))
        u_ref = torch.stack(u_ref)
        l_ref = c0.Re * 1e-3 / u_ref
        u_ref = u_ref[f0.batch].unsqueeze(-1)
        l_ref = l_ref[f0.batch].unsqueeze(-1)
        p_ref, d_ref = u_ref ** 2, u_ref * l_ref

        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        _, _, raw = self.encode_process_decode(c_graph.x, f_graph.x, get_topology(graphs))
        edge_attr_out = torch.cat([raw[:, 0:1] * u_ref, raw[:, 1:2] * u_ref * self.anisotropy_ratio,
                                   raw[:, 2:3] * p_ref, raw[:, 3:5] * d_ref], dim=-1)
        self.dt = c_graph.dt.clone()
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode != "rollout":
            output = self.normalizer.output(output)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def loss(self, output, graphs):
        return _real_space_loss(self, output, graphs)

    class Integrator(nn.Module):   # Fvgn.py:1380-1416 (one diffusion column, broadcast over both components)
        def __init__(self, config, rho, nu=None):
            super().__init__()
            self.rho, self.nu = rho, nu

        def forward(self, edge_output, c_graph, f_graph, dt):
            cf, d = f_graph.face, edge_output[:, 3:4]
            phi_a, phi_p = _advection_pressure(edge_output, c_graph, f_graph)
            phi_d = d[cf[0], :] + d[cf[1], :] + d[cf[2], :]
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p + phi_d * 1e-3)
