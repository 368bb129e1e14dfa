"""Domain decomposition of one large mesh for the multi-GPU processor (BASELINE.json config 4).

Pure host-side index work (torch CPU tensors): no CUDA, no torch.distributed - so the plan is testable on
CPU and identical on every rank (each rank computes the whole plan deterministically and keeps its part).

The reference has no counterpart (it runs one mesh on one GPU).  The decomposition follows from the
data-flow of its GN_Block (Fvgn.py:274-325, Mgn.py:216-267):

* a rank OWNS a contiguous strip of cells (equal counts along the sort key, e.g. the centroid x);
* its LOCAL FACES are all faces incident to a vertex of an owned cell - exactly the faces whose latents
  enter the two-hop aggregation of an owned cell; their latents are updated redundantly, never sent;
* its GHOST CELLS are the vertex-star of the owned cells (every cell sharing a vertex with an owned cell):
  the cells whose latent the face MLP of a local face gathers;
* ONE exchange of ghost-cell latents per GN_Block: the block-input ``x`` in MGN order, the raw cell-MLP
  output ``x'`` in FVGN order.

Local orderings keep global order (cells: owned ascending, then ghosts grouped by owner rank ascending;
faces and vertices ascending), so the receiver-sorted CSR of a part visits a vertex's contributions in the
same relative order as the single-GPU CSR and the partitioned result is bit-identical on owned rows.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass
class Partition:
    rank: int
    world: int
    n_owned: int
    cells: torch.Tensor            # [n_local] global cell ids: owned first, then ghosts by (owner, id)
    faces: torch.Tensor            # [E_loc] global face ids, ascending
    verts: torch.Tensor            # [V_loc] global vertex ids, ascending
    c_edge_index: torch.Tensor     # [2, E_loc] local cell ids
    v_edge_index: torch.Tensor     # [2, E_loc] local vertex ids
    v_face: torch.Tensor           # [3, n_owned] local vertex ids of the owned cells
    f_face: torch.Tensor           # [3, n_owned] local face ids of the owned cells (or empty)
    recv: Dict[int, Tuple[int, int]] = field(default_factory=dict)   # peer -> (first ghost row, count)
    send: Dict[int, torch.Tensor] = field(default_factory=dict)      # peer -> local owned rows to send (int64)

    @property
    def n_local(self) -> int:
        return int(self.cells.numel())

    @property
    def n_ghost(self) -> int:
        return self.n_local - self.n_owned


def assign_owners(key: torch.Tensor, world: int) -> torch.Tensor:
    """Equal-count strips along ``key`` (stable): owner rank per cell."""
    n = key.numel()
    order = torch.argsort(key, stable=True)
    owner = torch.empty(n, dtype=torch.int64)
    bounds = [(n * r) // world for r in range(world + 1)]
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def partition_mesh(c_edge_index: torch.Tensor, v_edge_index: torch.Tensor, v_face: torch.Tensor, key: torch.Tensor,
                   world: int, f_face: torch.Tensor | None = None) -> List[Partition]:
    """All ``world`` partitions of the mesh (see module docstring).  Inputs are the reference's global index
    tensors: ``c_graph.edge_index`` [2,E], ``v_graph.edge_index`` [2,E], ``v_graph.face`` [3,N], optional
    ``f_graph.face`` [3,N]."""
    c_ei, v_ei, vf = c_edge_index.cpu().long(), v_edge_index.cpu().long(), v_face.cpu().long()
    n_cells, n_verts = vf.shape[1], int(max(int(v_ei.max()), int(vf.max()))) + 1
    n_faces = c_ei.shape[1]
    owner = assign_owners(key.cpu(), world)
    parts: List[Partition] = []
    for r in range(world):
        owned = torch.nonzero(owner == r).flatten()                      # ascending
        vmask = torch.zeros(n_verts, dtype=torch.bool)
        vmask[vf[:, owned].reshape(-1)] = True
        fsel = vmask[v_ei[0]] | vmask[v_ei[1]]
        faces = torch.nonzero(fsel).flatten()
        cmask = torch.zeros(n_cells, dtype=torch.bool)
        cmask[c_ei[:, faces].reshape(-1)] = True
        cmask[owned] = False
        ghosts = torch.nonzero(cmask).flatten()
        gkey = owner[ghosts] * n_cells + ghosts
        ghosts = ghosts[torch.argsort(gkey)]
        cells = torch.cat([owned, ghosts])
        cmap = torch.full((n_cells,), -1, dtype=torch.int64)
        cmap[cells] = torch.arange(cells.numel())
        vsel = torch.zeros(n_verts, dtype=torch.bool)
        vsel[v_ei[:, faces].reshape(-1)] = True
        verts = torch.nonzero(vsel).flatten()
        vmap = torch.full((n_verts,), -1, dtype=torch.int64)
        vmap[verts] = torch.arange(verts.numel())
        fmap = None
        if f_face is not None:
            fmap = torch.full((n_faces,), -1, dtype=torch.int64)
            fmap[faces] = torch.arange(faces.numel())
        part = Partition(rank=r, world=world, n_owned=int(owned.numel()), cells=cells, faces=faces, verts=verts,
                         c_edge_index=cmap[c_ei[:, faces]], v_edge_index=vmap[v_ei[:, faces]],
                         v_face=vmap[vf[:, owned]],
                         f_face=fmap[f_face.cpu().long()[:, owned]] if f_face is not None else torch.empty(3, 0, dtype=torch.int64))
        assert int(part.c_edge_index.min()) >= 0 and int(part.v_face.min()) >= 0
        gown = owner[ghosts]
        start = part.n_owned
        for peer in range(world):
            cnt = int((gown == peer).sum())
            if cnt:
                part.recv[peer] = (start, cnt)
                start += cnt
        parts.append(part)
    # send lists: what each peer's ghost list asks of me, in the peer's ghost order
    for b in parts:
        for a_rank, (start, cnt) in b.recv.items():
            a = parts[a_rank]
            want = b.cells[start:start + cnt]                                 # global ids owned by a
            pos = torch.searchsorted(a.cells[:a.n_owned], want)
            assert torch.equal(a.cells[:a.n_owned][pos], want)
            a.send[b.rank] = pos
    return parts


def local_graphs(graphs, part: Partition):
    """Slice the reference's [c_graph, f_graph, v_graph] triplet to one partition (owned + ghost cells, local
    faces, local vertices) with local index tensors.  Per-cell attributes keep all local cells (ghosts are
    needed as encoder inputs); ``v_graph.face`` / ``f_graph.face`` cover the owned cells only."""
    from .graph import Data
    c, f, v = graphs
    cells, faces, verts = part.cells, part.faces, part.verts
    n_cells, n_faces = c.x.shape[0], c.edge_index.shape[1]

    def take(g, idx, n, skip):
        out = {}
        for k in g.keys():
            t = g[k]
            if k in skip or not torch.is_tensor(t):
                continue
            if t.dim() >= 1 and t.shape[0] == n:
                out[k] = t.cpu()[idx].contiguous()
            elif t.dim() == 0 or k == "dt":
                out[k] = t
        return out

    lc = Data(**take(c, cells, n_cells, ("edge_index", "face")))
    lc.edge_index = part.c_edge_index
    lf = Data(**take(f, faces, n_faces, ("edge_index", "face")))
    if part.f_face.numel():
        lf.face = part.f_face
    lv = Data(pos=v.pos.cpu()[verts].contiguous() if "pos" in v else torch.zeros(verts.numel(), 2))
    lv.edge_index = part.v_edge_index
    lv.face = part.v_face
    return [lc, lf, lv]
