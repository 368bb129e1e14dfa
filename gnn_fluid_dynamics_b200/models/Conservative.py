"""Conservative family on the B200 kernels - drop-in for reference ``src/models/Conservative.py``
(ConservativeA): symmetric + antisymmetric face encodings, sum-form face block, signed direct
edge->cell aggregation (Conservative.py:191-262).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import processor as P
from .._lib import ACT_TANH
from ..topology import get_topology
from .base import build_mlp, build_mlp_antisym, col, n_class_types
from .Fvgn import FvgnA
from .Mgn import MgnA


class ConservativeA(FvgnA):
    family = "cons_a"
    _registry_overrides = {
        "face_velocity_diff_char": lambda graphs: torch.norm(graphs[1].x_asym[:, 0:2], dim=1)}

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.encoder = self.Encoder(config, self.input_sizes, self.hidden_size)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])
        self.decoder = self.Decoder(config, self.hidden_size, self.output_sizes)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 3 + n_class_types(dataset), 0], [0, 5, 0])   # Conservative.py:62-64

    @classmethod
    def normalisation_tables(cls):   # Conservative.py:105-145
        z = "z_score"
        kinds = {k: z for k in ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x",
                                "cell_velocity_change_y", "face_area", "face_adjacent_distance",
                                "face_velocity_x", "face_velocity_y", "face_pressure"]}
        kinds["face_velocity_diff_char"] = "mean_scale"
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (1, "x_asym", col(0, 2), "face_velocity_diff_char"),
                  (1, "x_symm", col(0), "face_area"), (1, "x_symm", col(2), "face_adjacent_distance"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (1, "y", col(0), "face_velocity_x"), (1, "y", col(1), "face_velocity_y"),
                  (1, "y", col(2), "face_pressure")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (1, col(0), "face_velocity_x"), (1, col(1), "face_velocity_y"),
                   (1, col(2), "face_pressure")]
        return kinds, inputs, outputs

    def encode_process_decode(self, c_x, f_x_symm, f_x_asym, topo, hook=None):
        prec = self.prec
        e = P.mlp_rows(self.encoder.faceS_mlp, f_x_symm, prec)
        e_asym = P.mlp_rows(self.encoder.faceA_mlp, f_x_asym, prec, act=ACT_TANH)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        # GN_Block returns a fresh Data without edge_attr_asym, so the asym multiply fires in block 0
        # only (Conservative.py:220, 232-233); run_processor reproduces that.
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, e_asym=e_asym, hook=hook)
        return x, e, P.mlp_rows(self.decoder.face_mlp, e, prec)

    @staticmethod
    def _attach_signed_ell(topo, f_graph):
        """Fixed-degree table of the signed edge->cell aggregation (``MeshTopology.build_signed_cell_ell``) for the fused
        inference path of the 'cons_a' GN_Block; triangle meshes only (``f_graph.face`` [3, N])."""
        face = getattr(f_graph, "face", None) if P.FUSE_SIGNED_SUM else None      # (off by default: measured slower)
        topo.signed_ell = (topo.build_signed_cell_ell(face)
                           if torch.is_tensor(face) and face.dim() == 2 and face.shape[0] == 3 and face.shape[1] == topo.n_cells
                           else None)

    def forward(self, graphs, mode="rollout"):   # Conservative.py:164-189
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x_symm
        c_graph.edge_attr_asym = f_graph.x_asym
        topo = get_topology(graphs, need_cell_csr=True, two_hop=False)
        self._attach_signed_ell(topo, f_graph)
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x_symm, f_graph.x_asym, topo)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def update_features(self, output, input_graphs):   # Conservative.py:147-162
        c_graph, f_graph, v_graph = input_graphs
        c_graph.x = output["cell_velocity"].detach()
        u = c_graph.x[:, :2]
        dv = u[c_graph.edge_index[0]] - u[c_graph.edge_index[1]]
        mask = ((f_graph.type == 2) | (f_graph.type == 1)).squeeze(-1)
        dv = torch.where(mask.unsqueeze(-1), f_graph.y[:, 0:2], dv)   # == dv[mask] = y[mask], without the host sync
        f_graph.x_asym[:, 0:2] = dv
        return [c_graph, f_graph, v_graph]

    class Encoder(nn.Module):   # Conservative.py:191-202
        def __init__(self, config, input_sizes, hidden_size):
            super().__init__()
            self.faceA_mlp = build_mlp_antisym(config, 4, hidden_size, hidden_size)
            self.faceS_mlp = build_mlp(config, input_sizes[1], hidden_size, hidden_size)
            self.cell_mlp = build_mlp(config, input_sizes[0], hidden_size, hidden_size)

    class GN_Block(nn.Module):   # Conservative.py:204-254
        family = "cons_a"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block = self.Face_Block(config, hidden_size)
            self.cell_block = self.Cell_Block(config, hidden_size)

        class Face_Block(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp(config, hidden_size * 2, hidden_size, hidden_size)

        class Cell_Block(nn.Module):
            def __init__(self, config, hidden_size, mp_times=2):
                super().__init__()
                self.cell_mlp = build_mlp(config, hidden_size * 2, hidden_size, hidden_size)
                self.mp_times = mp_times


class ConservativeE(FvgnA):
    """Reference ``ConservativeE`` (Conservative.py:660-732): FvgnA encoder / decoder / integrator with
    face-block-first GN_Blocks whose cell block aggregates the raw face output directly onto cells - first half of
    the latent with equal signs ("symmetric"), second half with opposite signs ("antisymmetric")."""
    family = "cons_e"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])

    class GN_Block(nn.Module):
        family = "cons_e"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block = ConservativeA.GN_Block.Face_Block(config, hidden_size)    # in = 2H (e, x[row] + x[col])
            self.cell_block = ConservativeA.GN_Block.Cell_Block(config, hidden_size)    # in = 2H (x, sym | asym sums)


class ConservativeF(FvgnA):
    """Reference ``ConservativeF`` (Conservative.py:734-821): FVGN-order blocks whose cell block takes the symmetric
    half of the face latent through the vertices (two-hop mean) and the antisymmetric half as a signed direct
    edge->cell sum; concat-form face block."""
    family = "cons_f"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])

    class GN_Block(nn.Module):
        family = "cons_f"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.cell_block = ConservativeA.GN_Block.Cell_Block(config, hidden_size)    # in = 2H (x, two-hop sym, signed asym)
            self.face_block = FvgnA.GN_Block.Face_Block(config, hidden_size)            # in = 3H


class ConservativeB(MgnA):
    """ConservativeA's encoder and GN_Blocks under MgnA's node decoder, loss and outputs (Conservative.py:265-414)."""
    family = "cons_a"
    _registry_overrides = ConservativeA._registry_overrides

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)      # decoder = self.Decoder (node_mlp), Mgn.py:55
        self.encoder = ConservativeA.Encoder(config, self.input_sizes, self.hidden_size)
        self.processer_list = nn.ModuleList(
            [ConservativeA.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 3 + n_class_types(dataset), 0], [3, 0, 0])

    @classmethod
    def normalisation_tables(cls):   # Conservative.py:325-366
        z = "z_score"
        kinds = {k: z for k in ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
                                "cell_pressure", "face_area", "face_adjacent_distance", "face_velocity_x",
                                "face_velocity_y"]}
        kinds["face_velocity_diff_char"] = "mean_scale"
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (1, "x_asym", col(0, 2), "face_velocity_diff_char"),
                  (1, "x_symm", col(0), "face_area"), (1, "x_symm", col(2), "face_adjacent_distance"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (0, "y", col(2), "cell_pressure"),
                  (1, "y", col(0), "face_velocity_x"), (1, "y", col(1), "face_velocity_y")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (0, col(2), "cell_pressure")]
        return kinds, inputs, outputs

    def training_plan(self):
        raise NotImplementedError("ConservativeB trains through the per-op autograd wrappers (autograd_ops.py)")

    def encode_process_decode(self, c_x, f_x_symm, f_x_asym, topo, hook=None):
        prec = self.prec
        e = P.mlp_rows(self.encoder.faceS_mlp, f_x_symm, prec)
        e_asym = P.mlp_rows(self.encoder.faceA_mlp, f_x_asym, prec, act=ACT_TANH)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, e_asym=e_asym, hook=hook)
        return x, e, P.mlp_rows(self.decoder.node_mlp, x, prec)

    def forward(self, graphs, mode="rollout"):   # Conservative.py:383-404
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x_symm
        c_graph.edge_attr_asym = f_graph.x_asym
        topo = get_topology(graphs, need_cell_csr=True, two_hop=False)
        ConservativeA._attach_signed_ell(topo, f_graph)
        _, _, cell_output = self.encode_process_decode(c_graph.x, f_graph.x_symm, f_graph.x_asym, topo)
        output = [cell_output, None, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "cell_pressure": output[0][:, 2:3]}

    update_features = ConservativeA.update_features   # Conservative.py:367-381 (same body as A's)

    class Decoder(nn.Module):   # Conservative.py:406-414
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.node_mlp = build_mlp(config, hidden_size, hidden_size, output_sizes[0], norm_layer=False)


class ConservativeD(ConservativeA):
    """Reference ``ConservativeD`` (Conservative.py:417-658): same features / normalisation / encoder containers as
    ConservativeA, but the antisymmetric face encoding is a second latent stream with its own (bias-free tanh) face
    block on ``x[row] - x[col]``, the cell block aggregates both streams, and the decoder is
    ``final_mlp(symm_mlp(e_s) + asym_mlp(e_a))`` with antisymmetric asym / final heads."""
    family = "cons_d"

    def encode_process_decode(self, c_x, f_x_symm, f_x_asym, topo, hook=None):
        prec = self.prec
        e_s = P.mlp_rows(self.encoder.faceS_mlp, f_x_symm, prec)
        e_a = P.mlp_rows(self.encoder.faceA_mlp, f_x_asym, prec, act=ACT_TANH)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        # inference: the encoder's fresh outputs are advanced in place (TMA-store epilogue, processor.fast_mode)
        inplace = x.shape[0] > 0 and P.fast_mode(self.processer_list, [x, e_s, e_a], prec)
        for i, blk in enumerate(self.processer_list):
            x, e_s, e_a = P.gn_block_dual(blk, x, e_s, e_a, topo, prec, inplace=inplace)
            if hook is not None:
                hook(i, x.clone(), e_s.clone()) if inplace else hook(i, x, e_s)
        # decoder: symm head, asym head accumulated onto it through the residual epilogue, final antisymmetric head
        d_s = P.mlp_rows(self.decoder.symm_mlp, e_s, prec)
        _, comb = P.A.mlp(self.decoder.asym_mlp, [P.Seg(e_a)], e_a.shape[0], prec, act=ACT_TANH, residual=d_s,
                          want_raw=False, want_sum=True)
        self._last_e_asym = e_a
        return x, e_s, P.mlp_rows(self.decoder.final_mlp, comb, prec, act=ACT_TANH)

    class GN_Block(nn.Module):   # Conservative.py:572-645
        family = "cons_d"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block_symm = self.Face_Block_Symm(config, hidden_size)
            self.face_block_asym = self.Face_Block_Asym(config, hidden_size)
            self.cell_block = self.Cell_Block(config, hidden_size)

        class Face_Block_Symm(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp(config, hidden_size * 2, hidden_size, hidden_size)

        class Face_Block_Asym(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp_antisym(config, hidden_size * 2, hidden_size, hidden_size)

        class Cell_Block(nn.Module):
            def __init__(self, config, hidden_size, mp_times=2):
                super().__init__()
                self.cell_mlp = build_mlp(config, hidden_size * 3, hidden_size, hidden_size)
                self.mp_times = mp_times

    class Decoder(nn.Module):   # Conservative.py:647-658
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.symm_mlp = build_mlp(config, hidden_size, hidden_size, hidden_size, norm_layer=False)
            self.asym_mlp = build_mlp_antisym(config, hidden_size, hidden_size, hidden_size)
            self.final_mlp = build_mlp_antisym(config, hidden_size, hidden_size, output_sizes[1])


class ConservativeG(FvgnA):
    """Reference ``ConservativeG`` (Conservative.py:824-896): ConservativeF's hybrid cell block followed by the SUM-form
    face block ``face_mlp(cat[e, x'[row] + x'[col]])``."""
    family = "cons_g"

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.processer_list = nn.ModuleList(
            [self.GN_Block(config, self.hidden_size) for _ in range(config.model.mp_num)])

    class GN_Block(nn.Module):
        family = "cons_g"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.cell_block = ConservativeA.GN_Block.Cell_Block(config, hidden_size)    # in = 2H
            self.face_block = ConservativeA.GN_Block.Face_Block(config, hidden_size)    # in = 2H (e, x'[row] + x'[col])


class ConservativeI(ConservativeG):
    """Reference ``ConservativeI`` (Conservative.py:1211-1317): ConservativeG whose GN_Blocks re-impose the boundary
    conditions - the latent of INFLOW / WALL_BOUNDARY faces is reset to its block input after the residual add."""
    family = "cons_i"

    def forward_normalised(self, graphs, mode="rollout"):
        f_type = graphs[1].type
        keep = ~((f_type == 2) | (f_type == 1)).reshape(-1)                 # INFLOW = 2, WALL_BOUNDARY = 1 (OpenFoam.py:19-24)
        self._e_keep = keep.to(torch.float32).unsqueeze(1).expand(-1, self.hidden_size).contiguous()
        return super().forward_normalised(graphs, mode)

    def encode_process_decode(self, c_x, f_x, topo, hook=None, e_keep=None):
        prec = self.prec
        e_keep = e_keep if e_keep is not None else self._e_keep
        e = P.mlp_rows(self.encoder.face_mlp, f_x, prec)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        x, e, _ = P.run_processor(self.family, self.processer_list, x, e, topo, prec, hook=hook, e_keep=e_keep)
        return x, e, P.mlp_rows(self.decoder.face_mlp, e, prec)

    class GN_Block(ConservativeG.GN_Block):
        family = "cons_i"


class ConservativeH(ConservativeD):
    """Reference ``ConservativeH`` (Conservative.py:899-1208): symmetric / antisymmetric separation throughout -
    6 symmetric face features (area, one-hot type), std-scaled antisymmetric features, cell-block-first dual-stream
    GN_Blocks with the symmetric stream aggregated through the vertices, even / odd decoder heads
    (``q_n = softplus(even) * tanh(odd)``) and a normal-flux diffusion term in the integrator."""
    family = "cons_h"
    _registry_overrides = {}

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1)

    @classmethod
    def get_feature_sizes(cls, dataset):
        return ([2, 1 + n_class_types(dataset), 0], [0, 5, 0])   # Conservative.py:1012-1013

    @classmethod
    def normalisation_tables(cls):   # Conservative.py:948-996
        z, sd = "z_score", "std_scale"
        kinds = {k: z for k in ["cell_velocity_x", "cell_velocity_y", "cell_velocity_change_x", "cell_velocity_change_y",
                                "face_area", "face_velocity_x", "face_velocity_y", "face_pressure"]}
        kinds.update({k: sd for k in ["face_velocity_diff_x", "face_velocity_diff_y", "face_edge_vector_x",
                                      "face_edge_vector_y"]})
        inputs = [(0, "x", col(0), "cell_velocity_x"), (0, "x", col(1), "cell_velocity_y"),
                  (1, "x_asym", col(0), "face_velocity_diff_x"), (1, "x_asym", col(1), "face_velocity_diff_y"),
                  (1, "x_symm", col(0), "face_area"),
                  (1, "x_asym", col(2), "face_edge_vector_x"), (1, "x_asym", col(3), "face_edge_vector_y"),
                  (0, "y", col(0), "cell_velocity_change_x"), (0, "y", col(1), "cell_velocity_change_y"),
                  (1, "y", col(0), "face_velocity_x"), (1, "y", col(1), "face_velocity_y"),
                  (1, "y", col(2), "face_pressure")]
        outputs = [(0, col(0), "cell_velocity_change_x"), (0, col(1), "cell_velocity_change_y"),
                   (1, col(0), "face_velocity_x"), (1, col(1), "face_velocity_y"), (1, col(2), "face_pressure")]
        return kinds, inputs, outputs

    def forward(self, graphs, mode="rollout"):   # Conservative.py:1015-1036: two-hop topology + cell CSR
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x_symm
        c_graph.edge_attr_asym = f_graph.x_asym
        topo = get_topology(graphs, need_cell_csr=True, two_hop=True)
        _, _, edge_attr_out = self.encode_process_decode(c_graph.x, f_graph.x_symm, f_graph.x_asym, topo)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode == "rollout":
            output = self.normalizer.output(output, inverse=True)
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def encode_process_decode(self, c_x, f_x_symm, f_x_asym, topo, hook=None):
        prec = self.prec
        e_s = P.mlp_rows(self.encoder.faceS_mlp, f_x_symm, prec)
        e_a = P.mlp_rows(self.encoder.faceA_mlp, f_x_asym, prec, act=ACT_TANH)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        inplace = x.shape[0] > 0 and P.fast_mode(self.processer_list, [x, e_s, e_a], prec)      # see ConservativeD
        for i, blk in enumerate(self.processer_list):
            x, e_s, e_a = P.gn_block_dual_two_hop(blk, x, e_s, e_a, topo, prec, inplace=inplace)
            if hook is not None:
                hook(i, x.clone(), e_s.clone()) if inplace else hook(i, x, e_s)
        self._last_e_asym = e_a
        # decoder (Conservative.py:1186-1208): even head on cat[h+, h-^2], odd head on cat[h-, h+]
        n_e = e_s.shape[0]
        even, _ = P.A.mlp(self.decoder.even_mlp, [P.Seg(e_s), P.Seg(e_a * e_a)], n_e, prec)
        odd, _ = P.A.mlp(self.decoder.odd_mlp, [P.Seg(e_a), P.Seg(e_s)], n_e, prec, act=ACT_TANH)
        q_n = torch.nn.functional.softplus(even[:, 3:5]) * torch.tanh(odd)
        return x, e_s, torch.cat([even[:, 0:3], q_n], dim=-1)

    class Integrator(nn.Module):   # Conservative.py:1038-1083: diffusion term = signed normal flux q . n . area
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.face_area_norm = nn.BatchNorm1d(1)
            self.face_area = None

        def forward(self, edge_output, c_graph, f_graph, dt):
            from .Fvgn import flux_dot, normalize_face_area
            unv, cf = c_graph.normal, f_graph.face
            area = normalize_face_area(f_graph.area, c_graph.volume, c_graph.edge_index, dt, self.face_area_norm)
            self.face_area = area
            uv, p_face, q_face = edge_output[:, :2], edge_output[:, 2:3], edge_output[:, 3:]
            uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)
            phi_a = sum(flux_dot(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
            phi_d = sum(q_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return 1.0 * (-phi_a - phi_p / self.rho) + phi_d

    class GN_Block(ConservativeD.GN_Block):   # same containers, cell block first (Conservative.py:1098-1127)
        family = "cons_h"

    class Decoder(nn.Module):   # Conservative.py:1186-1208
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.even_mlp = build_mlp(config, 2 * hidden_size, hidden_size, 5, norm_layer=False)
            self.odd_mlp = build_mlp_antisym(config, 2 * hidden_size, hidden_size, 2)


class ConservativeJ(ConservativeH):
    """Reference ``ConservativeJ`` (Conservative.py:1320-1683): ConservativeH's encoder / dual-stream blocks / even-odd
    decoder, with learnt real-space output scales and a physical (un-normalised) integrator, as FvgnJ does for FvgnA."""

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__(config, loss_func, dataset, stats)
        self.integrator = self.Integrator(config, rho=1.0)
        self.velocity_scale_x = nn.Parameter(torch.tensor(1.0))
        self.velocity_scale_y = nn.Parameter(torch.tensor(0.01))
        self.pressure_scale = nn.Parameter(torch.tensor(1.0))
        self.diffusion_scale = nn.Parameter(torch.tensor(1.0))
        self.velocity_bias_x = nn.Parameter(torch.tensor(0.0))
        self.velocity_bias_y = nn.Parameter(torch.tensor(0.0))
        self.pressure_bias = nn.Parameter(torch.tensor(0.0))

    def forward(self, graphs, mode="rollout"):   # Conservative.py:1483-1517
        graphs = self.normalizer.input(graphs)
        c_graph, f_graph, v_graph = graphs
        c_graph.edge_attr = f_graph.x_symm
        c_graph.edge_attr_asym = f_graph.x_asym
        topo = get_topology(graphs, need_cell_csr=True, two_hop=True)
        _, _, raw = self.encode_process_decode(c_graph.x, f_graph.x_symm, f_graph.x_asym, topo)
        edge_attr_out = torch.cat([raw[:, 0:1] * self.velocity_scale_x + self.velocity_bias_x,
                                   raw[:, 1:2] * self.velocity_scale_y + self.velocity_bias_y,
                                   raw[:, 2:3] * self.pressure_scale + self.pressure_bias,
                                   raw[:, 3:5] * self.diffusion_scale], dim=-1)
        self.dt = c_graph.dt
        acc_pred = self.integrator(edge_attr_out, c_graph, f_graph, self.dt)
        output = [acc_pred, edge_attr_out, None]
        if mode != "rollout":
            output = self.normalizer.output(output)      # normalised for the training loss
        return {"cell_velocity_change": output[0][:, 0:2], "face_velocity": output[1][:, :2],
                "face_pressure": output[1][:, 2:3]}

    def loss(self, output, graphs):   # Conservative.py:1442-1477: continuity with the normalised face-area feature
        from .Fvgn import flux_dot
        c_graph, f_graph, v_graph = graphs
        lf = self.mse_term
        ff, unv, fv, area = f_graph.face, c_graph.normal, output["face_velocity"], f_graph.x_symm[:, 0:1]
        div = sum(flux_dot(fv[ff[j]], unv[:, j, :]) * area[ff[j]] for j in range(3))
        continuity = lf(div, torch.zeros_like(div), None, c_graph.batch)
        cvc = lf(output["cell_velocity_change"], c_graph.y, None, c_graph.batch)
        fvl = lf(output["face_velocity"], f_graph.y[:, :2], ~f_graph.boundary_mask, f_graph.batch)
        fpl = lf(output["face_pressure"], f_graph.y[:, 2:3], None, f_graph.batch)
        w = self.config.training.loss_weights
        total = (w["continuity"] * continuity + w["cell_velocity_change"] * cvc
                 + w["face_velocity"] * fvl + w["face_pressure"] * fpl)
        return {"total_log_loss": torch.mean(torch.log(total)), "continuity_loss": continuity,
                "cell_velocity_change_loss": cvc, "face_velocity_loss": fvl, "face_pressure_loss": fpl}

    class Integrator(nn.Module):   # Conservative.py:1520-1557
        def __init__(self, config, rho):
            super().__init__()
            self.rho = rho
            self.nu = 0.001

        def forward(self, edge_output, c_graph, f_graph, dt):
            from .Fvgn import flux_dot
            unv, cf, area = c_graph.normal, f_graph.face, f_graph.area
            uv, p_face, q_face = edge_output[:, 0:2], edge_output[:, 2:3], edge_output[:, 3:5]
            uu_vu = torch.cat([uv[:, 0:1] * uv, uv[:, 1:2] * uv], dim=-1)
            phi_a = sum(flux_dot(uu_vu[cf[j]], unv[:, j, :]) * area[cf[j]] for j in range(3))
            phi_d = sum(q_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            phi_p = sum(p_face[cf[j]] * unv[:, j, :] * area[cf[j]] for j in range(3))
            return torch.mean(dt) / c_graph.volume * (-phi_a - phi_p / self.rho + self.nu * phi_d)


def _padded_head(seq, act, n_valid, h=128):
    """MLPWeights of a 3-Linear MLP whose last Linear has ``n_valid`` < 128 outputs, zero-padded to 128 output rows so
    the 128-wide kernel path runs it; columns >= n_valid of the result are exactly zero.  Cached per weights version."""
    w = P.weights_of(seq, act)
    cached = getattr(w, "_padded", None)
    if cached is None:
        w3 = torch.zeros(h, w.w3.shape[1], dtype=torch.float32, device=w.w3.device)
        w3[:n_valid] = w.w3
        b3 = None
        if w.b3 is not None:
            b3 = torch.zeros(h, dtype=torch.float32, device=w.w3.device)
            b3[:n_valid] = w.b3
        cached = P.MLPWeights(w1=w.w1, b1=w.b1, w2=w.w2, b2=w.b2, w3=w3, b3=b3, has_ln=False, act=act)
        w._padded = cached
    return cached


class ConservativeK(ConservativeH):
    """Reference ``ConservativeK`` (Conservative.py:1685-1954): ConservativeH with the antisymmetric edge stream at HALF
    the hidden width.  The 64-wide stream is carried as a [E, 128] matrix whose upper 64 columns are identically zero
    (its MLPs run the 128-wide kernel path with the last Linear zero-padded to 128 output rows), so every consumer
    simply reads a 64-column segment of it; forward / rollout only."""
    family = "cons_h"

    def encode_process_decode(self, c_x, f_x_symm, f_x_asym, topo, hook=None):
        prec, hh = self.prec, self.hidden_size // 2
        if self.wants_grad():
            raise NotImplementedError("ConservativeK runs forward / rollout only on the B200 path (wrap the call in torch.no_grad())")
        ops, Seg = P.ops, P.Seg
        e_s = P.mlp_rows(self.encoder.faceS_mlp, f_x_symm, prec)
        x = P.mlp_rows(self.encoder.cell_mlp, c_x, prec)
        e_a, _ = ops.mlp_forward([Seg(f_x_asym.contiguous())], _padded_head(self.encoder.faceA_mlp, ACT_TANH, hh),
                                 f_x_asym.shape[0], prec)                               # [E, 128], columns 64.. = 0
        off, perm = topo.build_cell_csr()
        for i, blk in enumerate(self.processer_list):
            vsum = ops.segment_sum(e_s, e_s, 0, 0, P.H, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices)
            asym = ops.segment_sum(e_a, e_a, 0, 0, hh, -1.0, off, perm, topo.n_cells)   # [N, 64]
            x_raw, x_new = ops.mlp_forward([Seg(x), Seg(vsum, P.SEG_MEAN3, topo.vf), Seg(asym)],
                                           P.weights_of(blk.cell_block.cell_mlp), x.shape[0], prec, residual=x,
                                           want_raw=True, want_sum=True)
            _, s_new = ops.mlp_forward([Seg(e_s), Seg(x_raw, P.SEG_SUM2, (topo.row, topo.col))],
                                       P.weights_of(blk.face_block_symm.face_mlp), e_s.shape[0], prec, residual=e_s,
                                       want_raw=False, want_sum=True)
            _, a_new = ops.mlp_forward([Seg(e_a, width=hh), Seg(x_raw, P.SEG_DIFF2, (topo.row, topo.col))],
                                       _padded_head(blk.face_block_asym.face_mlp, ACT_TANH, hh), e_a.shape[0], prec,
                                       residual=e_a, want_raw=False, want_sum=True)
            x, e_s, e_a = x_new, s_new, a_new
            if hook is not None:
                hook(i, x, e_s)
        self._last_e_asym = e_a[:, :hh]
        n_e = e_s.shape[0]
        ea_sq = (e_a[:, :hh] * e_a[:, :hh]).contiguous()
        even, _ = ops.mlp_forward([Seg(e_s), Seg(ea_sq)], P.weights_of(self.decoder.even_mlp), n_e, prec)
        odd, _ = ops.mlp_forward([Seg(e_a, width=hh), Seg(e_s)], P.weights_of(self.decoder.odd_mlp, ACT_TANH), n_e, prec)
        q_n = torch.nn.functional.softplus(even[:, 3:5]) * torch.tanh(odd)
        return x, e_s, torch.cat([even[:, 0:3], q_n], dim=-1)

    class Encoder(nn.Module):   # Conservative.py:1848-1860
        def __init__(self, config, input_sizes, hidden_size):
            super().__init__()
            self.faceA_mlp = build_mlp_antisym(config, 4, hidden_size, hidden_size // 2)
            self.faceS_mlp = build_mlp(config, input_sizes[1], hidden_size, hidden_size)
            self.cell_mlp = build_mlp(config, input_sizes[0], hidden_size, hidden_size)

    class GN_Block(nn.Module):   # Conservative.py:1862-1934
        family = "cons_h"

        def __init__(self, config, hidden_size):
            super().__init__()
            self.face_block_symm = ConservativeD.GN_Block.Face_Block_Symm(config, hidden_size)
            self.face_block_asym = self.Face_Block_Asym(config, hidden_size)
            self.cell_block = self.Cell_Block(config, hidden_size)

        class Face_Block_Asym(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.face_mlp = build_mlp_antisym(config, hidden_size // 2 + hidden_size, hidden_size, hidden_size // 2)

        class Cell_Block(nn.Module):
            def __init__(self, config, hidden_size):
                super().__init__()
                self.cell_mlp = build_mlp(config, 2 * hidden_size + hidden_size // 2, hidden_size, hidden_size)

    class Decoder(nn.Module):   # Conservative.py:1936-1954
        def __init__(self, config, hidden_size, output_sizes):
            super().__init__()
            self.even_mlp = build_mlp(config, hidden_size + hidden_size // 2, hidden_size, 5, norm_layer=False)
            self.odd_mlp = build_mlp_antisym(config, hidden_size + hidden_size // 2, hidden_size, 2)
