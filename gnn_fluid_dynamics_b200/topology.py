"""Device-resident mesh topology consumed by the kernels: int32 index tensors + receiver-sorted CSRs.

Built once per mesh (or per training batch) from the reference's graph objects:
``c_graph.edge_index`` (cells of each face), ``v_graph.edge_index`` (vertices of each face, same
column order), ``v_graph.face`` (vertices of each cell) - SURVEY.md Appendix A/B.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import choose_launch_overlap


def _tensor_key(t):
    """Identity of a tensor's contents as far as the host can tell without a sync: storage address + version counter
    (in-place writes bump it)."""
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape))


class MeshTopology:
    """``__gnnfd_shared__`` makes ``Data.clone()/to()`` pass the object by reference, so a topology
    attached to ``v_graph.topology`` survives the per-step ``g.clone()`` of the rollout loop
    (reference ``src/rollout.py:313``) without being rebuilt or copied."""

    __gnnfd_shared__ = True

    def __init__(self, c_edge_index: torch.Tensor, v_edge_index: Optional[torch.Tensor],
                 v_face: Optional[torch.Tensor], n_cells: int, n_vertices: Optional[int],
                 need_cell_csr: bool = False):
        dev = c_edge_index.device
        if dev.type != "cuda":
            raise RuntimeError("MeshTopology needs CUDA tensors (no CPU fallback on this path)")
        self.device = dev
        self.n_cells = int(n_cells)
        self.n_faces = int(c_edge_index.shape[1])
        self.n_vertices = None if n_vertices is None else int(n_vertices)
        cei = ops.index_narrow(c_edge_index, self.n_cells)
        self._flags = [cei._gnnfd_range_flag]
        self.row, self.col = cei[0], cei[1]          # contiguous views of the [2,E] tensor
        self.vtx_offsets = self.vtx_perm = None
        self.v0 = self.v1 = None
        self.vf = None
        if v_edge_index is not None:
            vei = ops.index_narrow(v_edge_index, self.n_vertices)
            self._flags.append(vei._gnnfd_range_flag)
            self.v0, self.v1 = vei[0], vei[1]
            # index vector cat[v_ei[0]; v_ei[1]] == the [2,E] tensor flattened row-major
            self.vtx_offsets, self.vtx_perm = ops.csr_build(vei.reshape(-1), self.n_vertices)
        if v_face is not None:
            vf = ops.index_narrow(v_face, self.n_vertices)
            self._flags.append(vf._gnnfd_range_flag)
            self.vf = (vf[0], vf[1], vf[2])
        self.cell_offsets = self.cell_perm = None
        self._cell_index = None
        if need_cell_csr:
            self.build_cell_csr()
        self._vtx_n_offsets = None
        self.static = False          # attach_topology: the caller vouches that the mesh does not change
        self._c_key = _tensor_key(c_edge_index)
        self._v_key = (_tensor_key(v_edge_index), _tensor_key(v_face))

    def refresh_orientation(self, c_edge_index: torch.Tensor):
        """New owner / neighbour orientation of the SAME mesh (the reference re-flips ``c_graph.edge_index`` per
        training sample, ``transforms.py:3-7``): row / col are re-narrowed and the cell CSRs (which depend on the
        orientation) are dropped and rebuilt on demand; the vertex CSR, ``vf`` and the vf-CSR are static per mesh and
        kept."""
        if c_edge_index.shape[1] != self.n_faces:
            raise RuntimeError("refresh_orientation: a different mesh (face count changed)")
        cei = ops.index_narrow(c_edge_index, self.n_cells)
        self._flags = self._flags[1:] + [cei._gnnfd_range_flag] if self._flags else [cei._gnnfd_range_flag]
        self.row, self.col = cei[0], cei[1]
        self.cell_offsets = self.cell_perm = None
        self._rowcol = self._rowcsr = self._colcsr = self._rc2csr = None
        self._c_key = _tensor_key(c_edge_index)
        return self

    def matches(self, c_edge_index, v_edge_index, v_face) -> str:
        """'same' (built from exactly these tensors, unmodified since), 'orientation' (same vertex-side tensors,
        another or a modified ``c_graph.edge_index``) or 'other'."""
        if (_tensor_key(v_edge_index), _tensor_key(v_face)) != self._v_key:
            return "other"
        return "same" if _tensor_key(c_edge_index) == self._c_key else "orientation"

    def build_cell_csr(self):
        """CSR of cat[col; row] over cells (Conservative.py:244-245)."""
        if self.cell_offsets is None:
            idx = torch.cat([self.col, self.row])
            self.cell_offsets, self.cell_perm = ops.csr_build(idx, self.n_cells)
        return self.cell_offsets, self.cell_perm

    def build_rowcol_csr(self):
        """CSR of cat[row; col] over cells: the transpose of the Face_Block gathers x[row], x[col]
        (Fvgn.py:294) - d x[n] = sum_{row(k)=n} dIn1[k] + sum_{col(k)=n} dIn2[k]."""
        if getattr(self, "_rowcol", None) is None:
            self._rowcol = ops.csr_build(torch.cat([self.row, self.col]), self.n_cells)
        return self._rowcol

    def build_row_csr(self):
        """CSR of ``row`` alone over cells (edge -> first cell): the node-side reduction of edge gradients."""
        if getattr(self, "_rowcsr", None) is None:
            self._rowcsr = ops.csr_build(self.row, self.n_cells)
        return self._rowcsr

    def build_row_col_interleaved_csr(self):
        """CSR over 2 N virtual rows: row 2 n = the faces whose first cell is n, row 2 n + 1 = the faces whose second cell
        is n (index vector cat[2 row; 2 col + 1]).  One segment sum over it reduces an [E, 128] matrix onto the cells as
        the [N, 256] matrix [S_row | S_col] (viewed [2 N, 128]) in ONE pass: both reductions of a cell's faces run next
        to each other, so every source row comes from DRAM once."""
        if getattr(self, "_rc2csr", None) is None:
            self._rc2csr = ops.csr_build(torch.cat([self.row * 2, self.col * 2 + 1]), 2 * self.n_cells)
        return self._rc2csr

    def build_signed_cell_ell(self, f_face: torch.Tensor):
        """The direct signed edge->cell aggregation of the Conservative models (``agg[c] = sum_{col(k)=c} e_k -
        sum_{row(k)=c} e_k``, Conservative.py:243-254) as a FIXED-degree table for ``SEG_SUM3S``: a triangle cell has
        exactly three faces (``f_graph.face`` [3, N]), each contributing +e (the cell is the face's second cell), -e (its
        first cell) or nothing (boundary face = self-loop: its +e and -e entries cancel).  The three slots of a cell are
        ordered like the reference's scatter_add visits them (all '+' entries in ascending face id, then the '-' entries),
        so an interior cell's fused sum is bit-identical to the sequential one.  -> (idx0, idx1, idx2) int32 [N] each,
        entries encoded as in include/gnnfd_b200.h (k, ~k, GNNFD_SUM3S_ZERO)."""
        key = (f_face.data_ptr(), f_face._version, self._c_key)
        cached = getattr(self, "_ell", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        from ._lib import SUM3S_ZERO
        cf = f_face.to(self.device).long()                        # [3, N] face ids
        if cf.shape[1] != self.n_cells:
            raise RuntimeError("build_signed_cell_ell: f_graph.face does not have one column per cell")
        cell = torch.arange(self.n_cells, device=self.device).unsqueeze(0).expand_as(cf)
        row, col = self.row.long()[cf], self.col.long()[cf]
        plus, minus = (col == cell) & (row != cell), (row == cell) & (col != cell)
        zero = ~(plus | minus)                                     # self-loops (and faces that do not touch the cell)
        enc = torch.where(plus, cf, torch.where(minus, -cf - 1, torch.full_like(cf, SUM3S_ZERO)))
        order = torch.where(zero, 2, torch.where(minus, 1, 0)) * (2 * self.n_faces + 2) + cf   # '+' by k, then '-' by k, zeros last
        perm = torch.argsort(order, dim=0, stable=True)
        enc = torch.gather(enc, 0, perm).to(torch.int32).contiguous()
        ell = (enc[0], enc[1], enc[2])
        self._ell = (key, ell)
        return ell

    def build_col_csr(self):
        if getattr(self, "_colcsr", None) is None:
            self._colcsr = ops.csr_build(self.col, self.n_cells)
        return self._colcsr

    def build_vf_csr(self):
        """CSR of cat[vf0; vf1; vf2] over vertices: the transpose of the 3-vertex mean (Fvgn.py:317-321)."""
        if getattr(self, "_vfcsr", None) is None:
            self._vfcsr = ops.csr_build(torch.cat(list(self.vf)), self.n_vertices)
        return self._vfcsr

    def vertex_csr_rows(self, n_rows: int):
        """Vertex CSR padded to ``n_rows`` >= V rows (Vertex_Block writes N rows, VertPot.py:221)."""
        if n_rows == self.n_vertices:
            return self.vtx_offsets
        if self._vtx_n_offsets is None or self._vtx_n_offsets.numel() != n_rows + 1:
            if n_rows < self.n_vertices:
                raise RuntimeError("vertex_csr_rows: n_rows < n_vertices")
            pad = self.vtx_offsets[-1:].expand(n_rows - self.n_vertices)
            self._vtx_n_offsets = torch.cat([self.vtx_offsets, pad]).contiguous()
        return self._vtx_n_offsets

    def validate(self):
        """Synchronising range check of every narrowed index tensor."""
        bad = sum(int(f.item()) for f in self._flags)
        if bad:
            raise RuntimeError("MeshTopology: an index is out of range for its node count")
        return self

    @classmethod
    def from_graphs(cls, graphs, need_cell_csr: bool = False, two_hop: bool = True) -> "MeshTopology":
        c, f, v = graphs
        n_cells = c.x.shape[0]
        if two_hop:
            n_vertices = v.num_nodes
            if n_vertices is None:
                raise RuntimeError("v_graph.num_nodes unavailable (needs pos or x)")
            return cls(c.edge_index, v.edge_index, v.face, n_cells, n_vertices, need_cell_csr)
        return cls(c.edge_index, None, None, n_cells, None, need_cell_csr)


def get_topology(graphs, need_cell_csr: bool = False, two_hop: bool = True) -> MeshTopology:
    """Topology attached by ``attach_topology`` if it matches the graphs, else a fresh build.  Every model forward starts
    here, so this is also where the launch policy of the pass is chosen (``_lib.choose_launch_overlap``)."""
    topo = _get_topology(graphs, need_cell_csr, two_hop)
    choose_launch_overlap(topo.n_faces, torch.is_grad_enabled())
    return topo


def _get_topology(graphs, need_cell_csr: bool, two_hop: bool) -> MeshTopology:
    c, _, v = graphs
    topo = getattr(v, "topology", None) if v is not None else None
    if topo is None:
        topo = getattr(c, "topology", None)
    if isinstance(topo, MeshTopology) and topo.n_faces == c.edge_index.shape[1] \
            and topo.n_cells == c.x.shape[0] and (not two_hop or topo.vtx_offsets is not None):
        # a topology hung on the graphs by attach_topology is static by contract (rollout: the per-step g.clone() of
        # rollout.py:313 gives every step fresh tensors of the same mesh).  Any other attached topology is only trusted
        # for the tensors it was built from: a re-flipped / modified c_graph.edge_index of the same mesh refreshes the
        # orientation-dependent part, anything else is rebuilt.
        how = "same" if topo.static else topo.matches(c.edge_index, v.edge_index if two_hop else None,
                                                      v.face if two_hop else None)
        if how == "orientation":
            topo.refresh_orientation(c.edge_index)
            how = "same"
        if how == "same":
            if need_cell_csr:
                topo.build_cell_csr()
            return topo
    return MeshTopology.from_graphs(graphs, need_cell_csr, two_hop)


def attach_topology(graphs, need_cell_csr: bool = False, two_hop: bool = True) -> MeshTopology:
    """Build the topology once and hang it on the graphs (static meshes: rollout, validation)."""
    topo = MeshTopology.from_graphs(graphs, need_cell_csr, two_hop).validate()
    topo.static = True
    graphs[0].topology = topo
    if graphs[2] is not None:
        graphs[2].topology = topo
    return topo
