"""``transform_features`` - the per-sample feature construction the reference's dataset calls on the model class
(``src/train.py:350``, ``src/rollout.py:286``, ``src/datasets/DataSet.py``): raw velocity / pressure / flux time series
-> network inputs ``x`` and targets ``y`` on the cell and face graphs.  Plain CPU tensor code on the data-loading side
of the hot path; restated here so the drop-in classes answer the whole classmethod interface.

One routine, parametrised by what differs between the reference's 14 implementations (cited per option); the helper
functions follow ``src/utils/transforms.py``.  Random numbers are drawn in the reference's order (noise first, then the
edge flip) with the same calls, so a seeded run reproduces the reference's sample bit for bit.
"""
from __future__ import annotations

import torch

NORMAL, WALL_BOUNDARY, INFLOW, OUTFLOW, SLIP = 0, 1, 2, 3, 4   # datasets/OpenFoam.py:19-24


def random_edge_flip(cell_edge_index):   # utils/transforms.py:3-7 (in place)
    rand = torch.randint(0, 2, (cell_edge_index.shape[1],), dtype=torch.bool)
    first, second = cell_edge_index[1, rand], cell_edge_index[0, rand]
    cell_edge_index[0, rand], cell_edge_index[1, rand] = first, second
    return cell_edge_index, rand


def add_noise(tensor, std):   # utils/transforms.py:19-22 (in place: the caller passes a view of the raw series)
    tensor += torch.normal(mean=0.0, std=std, size=tensor.shape)
    return tensor


def clean_graphs(graphs):   # utils/transforms.py:24-34
    c_graph, f_graph, v_graph = graphs
    del c_graph.velocity
    del c_graph.pressure
    del f_graph.velocity
    del f_graph.pressure
    del f_graph.flux
    return [c_graph, f_graph, v_graph]


def _n_types(dataset):
    return len(dataset.class_types)


def transform(dataset, graphs, *, cell_y, face_x, face_y, impose_bc=True, flip_flux=None, bundle=False, clean=True):
    """cell_y: 'change' | 'change+p' | 'target+p' | 'target' | 'bundle';  face_x: 'fvgn' | 'fvgn_h' | 'cons' | 'cons_h';
    face_y: 'uvp' | 'uv' | 'uvpf' | 'pf' | 'bundle';  flip_flux: None | 'all' | 'last'."""
    cell_graph, face_graph, vertex_graph = graphs
    cell_velocity = cell_graph.velocity[:, 0:1] if bundle else cell_graph.velocity[:, 0]
    if dataset.noise:
        cell_velocity = add_noise(cell_velocity, std=dataset.config.training.noise_std)
    cell_graph.x = cell_velocity.flatten(start_dim=1) if bundle else cell_velocity
    if cell_y == "change":            # Fvgn.py:108, Flux.py:68, Conservative.py:73
        cell_graph.y = cell_graph.velocity[:, -1] - cell_velocity
    elif cell_y == "change+p":        # Mgn.py:71-72, Conservative.py:289-291
        cell_graph.y = torch.cat([cell_graph.velocity[:, -1] - cell_velocity, cell_graph.pressure[:, -1]], dim=1)
    elif cell_y == "target+p":        # Mgn.py:294
        cell_graph.y = torch.cat([cell_graph.velocity[:, -1], cell_graph.pressure[:, -1]], dim=1)
    elif cell_y == "target":          # Fvgn.py:804
        cell_graph.y = cell_graph.velocity[:, -1]
    elif cell_y == "bundle":          # Fvgn.py:483-484
        cell_graph.y = cell_graph.velocity[:, 1:] - cell_velocity
    else:
        raise ValueError(cell_y)

    if dataset.mode == "train":       # random orientation flip; boundary self-loops keep their normal
        cell_graph.edge_index, flip_mask = random_edge_flip(cell_graph.edge_index)
        safe_flip = flip_mask & (cell_graph.edge_index[0] != cell_graph.edge_index[1])
        face_graph.normal[safe_flip] *= -1
        if flip_flux == "all":        # Flux.py:75
            face_graph.flux[safe_flip] *= -1
        elif flip_flux == "last":     # Flux.py:311
            face_graph.flux[:, -1][safe_flip] *= -1

    t = face_graph.type
    interior = (t == NORMAL) | (t == OUTFLOW) | (t == SLIP) | (t == WALL_BOUNDARY)
    face_graph.boundary_mask = ~interior.squeeze()

    u = cell_velocity.squeeze(1) if bundle else cell_velocity
    row, col = cell_graph.edge_index[0], cell_graph.edge_index[1]
    dv = u[row] - u[col]
    if impose_bc:                     # Fvgn.py:122 (FluxA / FluxC leave the difference as it is)
        dv[face_graph.boundary_mask] = face_graph.velocity[:, 0][face_graph.boundary_mask]
    edge_vec = cell_graph.pos[row] - cell_graph.pos[col]
    one_hot = torch.nn.functional.one_hot(t.squeeze(-1), num_classes=_n_types(dataset))

    if face_x == "fvgn":              # Fvgn.py:125
        face_graph.x = torch.cat([dv, edge_vec, face_graph.area, one_hot], dim=1)
    elif face_x == "fvgn_h":          # Fvgn.py:1044-1057
        dist = torch.norm(edge_vec, dim=1, keepdim=True)
        dot = torch.clamp(((edge_vec / (dist + 1e-8)) * face_graph.normal).sum(dim=1, keepdim=True), -1.0, 1.0)
        angle = torch.acos(torch.abs(dot))
        angle = torch.where(dist < 1e-8, torch.zeros_like(angle), angle)
        face_graph.x = torch.cat([dv, face_graph.normal, face_graph.area, dist, angle, one_hot], dim=1)
    elif face_x == "cons":            # Conservative.py:90-98
        ev_n = torch.nn.functional.normalize(edge_vec, dim=1)
        dist = torch.norm(edge_vec, dim=1, keepdim=True)
        n_n = torch.nn.functional.normalize(face_graph.normal, dim=1)
        angle = torch.acos(torch.clamp((ev_n * n_n).sum(dim=1, keepdim=True), -1.0, 1.0))
        face_graph.x_symm = torch.cat([face_graph.area, angle, dist, one_hot], dim=1)
        face_graph.x_asym = torch.cat([dv, n_n], dim=1)
    elif face_x == "cons_h":          # Conservative.py:940-941
        face_graph.x_symm = torch.cat([face_graph.area, one_hot], dim=1)
        face_graph.x_asym = torch.cat([dv, edge_vec], dim=1)
    else:
        raise ValueError(face_x)

    if face_y == "uvp":               # Fvgn.py:126
        face_graph.y = torch.cat([face_graph.velocity[:, -1], face_graph.pressure[:, -1]], dim=1)
    elif face_y == "uv":              # Mgn.py:91 (boundary values only)
        face_graph.y = face_graph.velocity[:, -1]
    elif face_y == "uvpf":            # Flux.py:85
        face_graph.y = torch.cat([face_graph.velocity[:, -1], face_graph.pressure[:, -1], face_graph.flux[:, -1]], dim=1)
    elif face_y == "pf":              # Flux.py:322
        face_graph.y = torch.cat([face_graph.pressure[:, -1], face_graph.flux[:, -1]], dim=1)
    elif face_y == "bundle":          # Fvgn.py:506-507
        face_graph.y = torch.cat([face_graph.velocity[:, 1:], face_graph.pressure[:, 1:]], dim=2)
    else:
        raise ValueError(face_y)

    if clean:
        return clean_graphs([cell_graph, face_graph, vertex_graph])
    return graphs


# which options each reference implementation uses; subclasses inherit their parent's entry exactly as in the reference
SPECS = {
    "FvgnA": dict(cell_y="change", face_x="fvgn", face_y="uvp"),                                       # Fvgn.py:101-131
    "FvgnC": dict(cell_y="bundle", face_x="fvgn", face_y="bundle", bundle=True),                       # Fvgn.py:476-510
    "FvgnD": dict(cell_y="target", face_x="fvgn", face_y="uvp", clean=False),                          # Fvgn.py:796-826
    "FvgnH": dict(cell_y="change", face_x="fvgn_h", face_y="uvp"),                                     # Fvgn.py:1022-1062
    "MgnA": dict(cell_y="change+p", face_x="fvgn", face_y="uv"),                                       # Mgn.py:63-96
    "MgnB": dict(cell_y="target+p", face_x="fvgn", face_y="uv"),                                       # Mgn.py:286-316
    "FluxA": dict(cell_y="change", face_x="fvgn", face_y="uvpf", impose_bc=False, flip_flux="all", clean=False),   # Flux.py:59-87
    "FluxC": dict(cell_y="change", face_x="fvgn", face_y="pf", impose_bc=False, flip_flux="last", clean=False),   # Flux.py:296-324
    "ConservativeA": dict(cell_y="change", face_x="cons", face_y="uvp"),                               # Conservative.py:66-103
    "ConservativeB": dict(cell_y="change+p", face_x="cons", face_y="uv"),                              # Conservative.py:281-322
    "ConservativeD": dict(cell_y="change", face_x="cons", face_y="uvp"),                               # Conservative.py:435-472
    "ConservativeH": dict(cell_y="change", face_x="cons_h", face_y="uvp"),                             # Conservative.py:915-946
    "ConservativeJ": dict(cell_y="change", face_x="cons_h", face_y="uvp"),                             # Conservative.py:1345-1375
    "ConservativeK": dict(cell_y="change", face_x="cons_h", face_y="uvp"),                             # Conservative.py:1704-1735
}


def transform_for(cls, dataset, graphs):
    """Resolve the spec through the class's MRO (the reference's inheritance decides which implementation runs)."""
    for klass in cls.__mro__:
        spec = SPECS.get(klass.__name__)
        if spec is not None:
            return transform(dataset, graphs, **spec)
    raise NotImplementedError(f"{cls.__name__}: no transform_features in the reference either")
