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


def normalize_face_area(face_area, cell_volume, edge_index, dt, batch_norm):
    """face_area * mean(dt) / mean adjacent cell volume, through BatchNorm1d(1)
    (reference utils/normalisation.py:325-344)."""
    vol = (cell_volume.index_select(0, edge_index[0]) + cell_volume.index_select(0, edge_index[1])) / 2
    return batch_norm((face_area * (torch.mean(dt) / vol)).view(-1, 1))


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
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook)
        return x, e, P.mlp_rows(self.decoder.face_mlp, e, prec)

    def forward(self, graphs, mode="rollout"):   # Fvgn.py:150-174
        return self.forward_normalised(self.normalizer.input(graphs), mode)

    def forward_normalised(self, graphs, mode="rollout"):
        """``forward`` after the in-place input normalisation (graphs already normalised and resident)."""
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x
        topo = get_topology(graphs)
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
        lf = self.loss_func
        face_area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, self.dt,
                                        self.integrator.face_area_norm)
        ff, unv, fv = f_graph.face, c_graph.normal, output["face_velocity"]
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
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
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
