"""Import shim that lets the reference's *model* code run in this container.

ONLY used here (where /root/reference exists) by ``make_golden.py`` to generate the committed
fixtures; nothing on the GPU box imports this.  The reference needs torch_geometric, torch_scatter,
h5py and pyvista, none of which is installed and none of which can be installed (no network), so
the four are replaced by minimal stand-ins with their documented semantics:

* ``torch_geometric.data.Data``  -> attribute bag (``gnn_fluid_dynamics_b200.graph.Data``)
* ``torch_scatter.scatter_add``  -> ``zeros(dim_size, F).index_add_(0, index, src)`` (dim=0 only)
* ``global_add_pool / global_mean_pool`` -> index_add_ / bincount
* ``h5py`` / ``pyvista``         -> empty modules (only touched by dataset I/O, which is not run)
"""
import os
import sys
import types

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF_SRC = "/root/reference/src"


def scatter_add(src, index, dim=0, dim_size=None, out=None):
    assert dim == 0, "stub covers the reference's call sites only (dim=0)"
    if dim_size is None:
        dim_size = int(index.max()) + 1
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


def _global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype).index_add_(0, batch, x)


def _global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    s = _global_add_pool(x, batch, size)
    cnt = torch.bincount(batch, minlength=size).clamp(min=1).to(x.dtype)
    return s / cnt.view(-1, *([1] * (x.dim() - 1)))


def install():
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from gnn_fluid_dynamics_b200.graph import Data

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dataset:  # torch_geometric.data.Dataset base used by src/datasets/DataSet.py
        def __init__(self, *a, **k):
            pass

    class _Loader:
        def __init__(self, *a, **k):
            raise RuntimeError("DataLoader is not available under the stub")

    tg = mod("torch_geometric")
    tg.data = mod("torch_geometric.data", Data=Data, Dataset=_Dataset, Batch=Data)
    mod("torch_geometric.data.dataset", Dataset=_Dataset)
    tg.loader = mod("torch_geometric.loader", DataLoader=_Loader)
    tg.nn = mod("torch_geometric.nn", global_add_pool=_global_add_pool,
                global_mean_pool=_global_mean_pool)
    tg.utils = mod("torch_geometric.utils", unbatch=None)
    tg.transforms = mod("torch_geometric.transforms")
    mod("torch_scatter", scatter_add=scatter_add)
    mod("h5py", File=None)
    mod("pyvista")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
