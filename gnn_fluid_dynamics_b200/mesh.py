"""Synthetic OpenFOAM-shaped triangular meshes and the reference's connectivity convention.

Test/bench infrastructure: there is no network and no dataset, so inputs are synthetic 2-D
triangular meshes with the shape of the reference's cases (``generate/mesh.py:276-302``: a
20a x 10a channel, a = 0.15, with an obstacle centred at (5a, 5a)).

``connectivity()`` is a vectorised restatement of the index conventions of
``src/utils/geometry.py:64-170`` (``compute_connectivity``) and ``:173-202`` (``reorder_face``): the
reference builds them with Python dict loops (hours at 4M cells); the conventions are normative
because every kernel consumes ``c_graph.edge_index``, ``v_graph.edge_index``, ``v_graph.face`` and
``f_graph.face`` in exactly that order.  ``tests/golden/connectivity_*.npz`` pins this function
against the reference routine run on the same inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .graph import Data

# reference src/datasets/OpenFoam.py:19-24
NODE_NORMAL, NODE_WALL, NODE_INFLOW, NODE_OUTFLOW, NODE_SLIP = 0, 1, 2, 3, 4
NUM_NODE_TYPES = 5


def connectivity(cells: np.ndarray, vertex_pos: np.ndarray):
    """(face_index[3,N], cell_edge_index[2,E], vertex_edge_index[2,E]), int64.

    * vertex edges are the unique rows of (max(vi,vj), min(vi,vj)) in lexicographic order; the
      position in that list is the face id (geometry.py:100-119);
    * ``face_index[j, c]`` is the face id of local edge j in ((v0,v1),(v1,v2),(v2,v0)) of cell c
      (geometry.py:128-136);
    * ``cell_edge_index[:, f]`` is (first, second) cell touching f in ascending cell id, a self-loop
      on the boundary (geometry.py:138-160), then oriented so that centroid[s]-centroid[r] has
      x > 0, or x == 0 and y > 0 (geometry.py:173-202).
    """
    cells = np.asarray(cells, dtype=np.int64)
    n = cells.shape[0]
    a = np.concatenate([cells[:, 0], cells[:, 1], cells[:, 2]])
    b = np.concatenate([cells[:, 1], cells[:, 2], cells[:, 0]])
    hi = np.maximum(a, b)
    lo = np.minimum(a, b)
    nv = int(vertex_pos.shape[0])
    key = hi * nv + lo
    uniq, inverse = np.unique(key, return_inverse=True)
    e = uniq.shape[0]
    vertex_edge_index = np.stack([uniq // nv, uniq % nv]).astype(np.int64)
    face_index = inverse.reshape(3, n).astype(np.int64)

    # first / second incident cell in the reference's visiting order (cell ascending, slot ascending)
    cell_of = np.tile(np.arange(n, dtype=np.int64), 3)
    slot_of = np.repeat(np.arange(3, dtype=np.int64), n)
    visit = np.lexsort((slot_of, cell_of))          # by cell, then slot
    f_sorted = inverse[visit]
    c_sorted = cell_of[visit]
    order = np.argsort(f_sorted, kind="stable")
    f2 = f_sorted[order]
    c2 = c_sorted[order]
    first_pos = np.searchsorted(f2, np.arange(e), side="left")
    last_pos = np.searchsorted(f2, np.arange(e), side="right") - 1
    first = c2[first_pos]
    second = c2[last_pos]                            # == first on the boundary
    cei = np.stack([first, second])

    centroids = vertex_pos[cells].mean(axis=1)
    vec = centroids[cei[0]] - centroids[cei[1]]
    keep = (vec[:, 0] > 0) | ((vec[:, 0] == 0) & (vec[:, 1] > 0))
    cell_edge_index = np.where(keep[None, :], cei, cei[::-1]).astype(np.int64)
    return face_index, cell_edge_index, vertex_edge_index


@dataclass
class Mesh:
    """Geometry + connectivity of one triangular mesh (numpy, host)."""
    cells: np.ndarray              # [N,3] vertex ids per cell
    vertex_pos: np.ndarray         # [V,2]
    face_index: np.ndarray         # [3,N]
    cell_edge_index: np.ndarray    # [2,E]
    vertex_edge_index: np.ndarray  # [2,E]
    face_type: np.ndarray          # [E,1]
    face_area: np.ndarray          # [E,1]
    face_normal: np.ndarray        # [E,2]
    face_pos: np.ndarray           # [E,2]
    cell_pos: np.ndarray           # [N,2]
    cell_volume: np.ndarray        # [N,1]
    cell_normal: np.ndarray        # [N,3,2]

    @property
    def n_cells(self):
        return self.cells.shape[0]

    @property
    def n_faces(self):
        return self.vertex_edge_index.shape[1]

    @property
    def n_vertices(self):
        return self.vertex_pos.shape[0]


def _inside_obstacle(p: np.ndarray, kind: str, a: float) -> np.ndarray:
    cx, cy = 5 * a, 5 * a
    if kind == "cylinder":
        return (p[:, 0] - cx) ** 2 + (p[:, 1] - cy) ** 2 < (0.5 * a) ** 2
    if kind == "ellipse":
        return ((p[:, 0] - cx) / (0.75 * a)) ** 2 + ((p[:, 1] - cy) / (0.4 * a)) ** 2 < 1.0
    if kind == "airfoil":  # NACA0012, chord 2a, leading edge at (4a, 5a)
        chord = 2 * a
        t = (p[:, 0] - (cx - a)) / chord
        ok = (t > 0) & (t < 1)
        tt = np.clip(t, 0, 1)
        half = 0.6 * chord * (0.2969 * np.sqrt(tt) - 0.1260 * tt - 0.3516 * tt ** 2
                              + 0.2843 * tt ** 3 - 0.1015 * tt ** 4)
        return ok & (np.abs(p[:, 1] - cy) < half)
    if kind == "none":
        return np.zeros(p.shape[0], dtype=bool)
    raise ValueError(f"unknown obstacle kind {kind!r}")


def make_mesh(n_cells: int, kind: str = "cylinder", seed: int = 0, jitter: float = 0.2,
              grading: float = 0.35, sort_cell_vertices: bool = False, a: float = 0.15) -> Mesh:
    """A jittered, graded triangulation of the 20a x 10a channel with an obstacle removed.

    ``kind``: cylinder | ellipse | airfoil | none.  About ``n_cells`` triangles (2 per quad of an
    nx x ny = 2:1 lattice, diagonal chosen at random so vertex degrees spread over 4..8 like an
    unstructured mesher's).  ``kind='none'`` with n_cells = 2*m*m gives exactly the
    2 048 / 20 000-cell structured cases quoted in SURVEY.md section 6.
    """
    rng = np.random.RandomState(seed)
    lx, ly = 20 * a, 10 * a
    if kind == "none":
        ny = max(1, int(round(math.sqrt(n_cells / 2))))
        nx = ny
    else:
        ny = max(2, int(round(math.sqrt(n_cells / 4))))
        nx = 2 * ny
    u = np.linspace(0.0, 1.0, nx + 1)
    v = np.linspace(0.0, 1.0, ny + 1)
    if kind != "none" and grading > 0:
        # cluster lattice rows around the obstacle height (v0 = 0.5) with a monotone cubic map
        v = 0.5 + (v - 0.5) * (1 - grading) + grading * 4 * (v - 0.5) ** 3
        v[0], v[-1] = 0.0, 1.0
    gx, gy = np.meshgrid(u * lx, v * ly, indexing="xy")       # [(ny+1),(nx+1)]
    pos = np.stack([gx.ravel(), gy.ravel()], axis=1)
    hx, hy = lx / nx, ly / ny
    li, lj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    interior = ((li > 0) & (li < nx) & (lj > 0) & (lj < ny)).ravel()
    if jitter > 0:
        hy_min = hy * ((1 - grading) if (kind != "none" and grading > 0) else 1.0)
        d = (rng.rand(pos.shape[0], 2) - 0.5) * 2 * jitter * np.array([hx, hy_min])
        pos = pos + d * interior[:, None]

    def vid(i, j):  # column i, row j
        return j * (nx + 1) + i

    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ii, jj = ii.ravel(), jj.ravel()
    v00, v10, v01, v11 = vid(ii, jj), vid(ii + 1, jj), vid(ii, jj + 1), vid(ii + 1, jj + 1)
    flip = rng.rand(ii.shape[0]) < 0.5
    # CCW triangles; two diagonal choices
    t1 = np.where(flip[:, None], np.stack([v00, v10, v01], 1), np.stack([v00, v10, v11], 1))
    t2 = np.where(flip[:, None], np.stack([v10, v11, v01], 1), np.stack([v00, v11, v01], 1))
    cells = np.empty((2 * ii.shape[0], 3), dtype=np.int64)
    cells[0::2] = t1
    cells[1::2] = t2

    if kind != "none":
        cen = pos[cells].mean(axis=1)
        keep = ~_inside_obstacle(cen, kind, a)
        cells = cells[keep]
        used = np.zeros(pos.shape[0], dtype=bool)
        used[cells.ravel()] = True
        remap = np.cumsum(used) - 1
        cells = remap[cells]
        pos = pos[used]
    if sort_cell_vertices:   # src/datasets/CylinderFlow.py:67 pre-sorts each cell's vertices
        cells = np.sort(cells, axis=1)
    return finalize_mesh(cells, pos, lx, ly)


def finalize_mesh(cells: np.ndarray, pos: np.ndarray, lx: float, ly: float) -> Mesh:
    face_index, cei, vei = connectivity(cells, pos)
    p0, p1 = pos[vei[0]], pos[vei[1]]
    evec = p1 - p0
    area = np.linalg.norm(evec, axis=1, keepdims=True)
    fnormal = np.stack([evec[:, 1], -evec[:, 0]], axis=1) / np.maximum(area, 1e-30)
    fpos = 0.5 * (p0 + p1)
    cpos = pos[cells].mean(axis=1)
    q0, q1, q2 = pos[cells[:, 0]], pos[cells[:, 1]], pos[cells[:, 2]]
    vol = 0.5 * np.abs((q1[:, 0] - q0[:, 0]) * (q2[:, 1] - q0[:, 1])
                       - (q1[:, 1] - q0[:, 1]) * (q2[:, 0] - q0[:, 0]))[:, None]
    # outward unit normal of each cell's 3 faces (geometry.py:205-268: flip if pointing to the centroid)
    cn = np.empty((cells.shape[0], 3, 2))
    for j in range(3):
        fn = fnormal[face_index[j]]
        inward = ((cpos - fpos[face_index[j]]) * fn).sum(axis=1) > 0
        cn[:, j, :] = np.where(inward[:, None], -fn, fn)
    boundary = cei[0] == cei[1]
    ftype = np.zeros((vei.shape[1], 1), dtype=np.int64)
    tol = 1e-9
    on_in = boundary & (np.abs(fpos[:, 0]) < tol)
    on_out = boundary & (np.abs(fpos[:, 0] - lx) < tol)
    ftype[boundary, 0] = NODE_WALL
    ftype[on_in, 0] = NODE_INFLOW
    ftype[on_out, 0] = NODE_OUTFLOW
    return Mesh(cells=cells, vertex_pos=pos, face_index=face_index, cell_edge_index=cei,
                vertex_edge_index=vei, face_type=ftype, face_area=area, face_normal=fnormal,
                face_pos=fpos, cell_pos=cpos, cell_volume=vol, cell_normal=cn)


def mesh_graphs(mesh: Mesh, seed: int = 0, flavour: str = "fvgn", dt: float = 0.01,
                flip_edges: bool = False):
    """The three graphs ``[c_graph, f_graph, v_graph]`` as the models' ``forward`` expects them
    *after* ``transform_features`` (reference ``Fvgn.py:101-131``; Conservative flavour
    ``Conservative.py:66-103``): synthetic N(0,1) features on the real connectivity.

    flavour 'fvgn': f.x = [du(2), dpos(2), area(1), one_hot(5)];  'conservative': f.x_symm[E,8],
    f.x_asym[E,4];  'conservative_h': f.x_symm[E,6] (area, one-hot), f.x_asym[E,4].  ``flip_edges`` applies the training-time random orientation flip
    (``utils/transforms.py:3-7``).
    """
    g = torch.Generator().manual_seed(seed)
    f32 = torch.float32
    n, e = mesh.n_cells, mesh.n_faces
    cei = torch.from_numpy(mesh.cell_edge_index.copy())
    fnormal = torch.from_numpy(mesh.face_normal).to(f32)
    if flip_edges:
        r = torch.randint(0, 2, (e,), generator=g, dtype=torch.bool)
        cei = torch.where(r[None, :], cei.flip(0), cei)
        safe = r & (cei[0] != cei[1])
        fnormal = torch.where(safe[:, None], -fnormal, fnormal)
    ftype = torch.from_numpy(mesh.face_type)
    one_hot = torch.nn.functional.one_hot(ftype.squeeze(-1), NUM_NODE_TYPES).to(f32)
    u = torch.randn(n, 2, generator=g)
    c = Data(x=u,
             y=torch.randn(n, 2, generator=g),
             pos=torch.from_numpy(mesh.cell_pos).to(f32),
             volume=torch.from_numpy(mesh.cell_volume).to(f32),
             normal=torch.from_numpy(mesh.cell_normal).to(f32),
             edge_index=cei,
             dt=torch.tensor(dt, dtype=f32))
    boundary_mask = (ftype.squeeze(-1) == NODE_INFLOW)
    f = Data(pos=torch.from_numpy(mesh.face_pos).to(f32),
             face=torch.from_numpy(mesh.face_index.copy()),
             type=ftype,
             area=torch.from_numpy(mesh.face_area).to(f32),
             normal=fnormal,
             boundary_mask=boundary_mask,
             y=torch.randn(e, 4, generator=g))
    area_n = torch.randn(e, 1, generator=g)
    if flavour == "fvgn":
        f.x = torch.cat([torch.randn(e, 4, generator=g), area_n, one_hot], dim=1)
    elif flavour == "conservative":
        f.x_symm = torch.cat([area_n, torch.rand(e, 1, generator=g) * 3.0,
                              torch.rand(e, 1, generator=g), one_hot], dim=1)
        f.x_asym = torch.cat([torch.randn(e, 2, generator=g),
                              torch.nn.functional.normalize(fnormal, dim=1)], dim=1)
    elif flavour == "conservative_h":      # ConservativeH features (Conservative.py:916-946)
        f.x_symm = torch.cat([area_n, one_hot], dim=1)
        f.x_asym = torch.randn(e, 4, generator=g)         # [du(2), cell edge vector(2)]
    else:
        raise ValueError(flavour)
    v = Data(pos=torch.from_numpy(mesh.vertex_pos).to(f32),
             edge_index=torch.from_numpy(mesh.vertex_edge_index.copy()),
             face=torch.from_numpy(mesh.cells.T.copy()))
    return [c, f, v]
