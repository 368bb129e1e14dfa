"""Graph containers for the mesh graphs the models consume.

The reference hands three PyG ``Data`` objects ``[c_graph, f_graph, v_graph]`` to every model
(reference ``src/datasets/DataSet.py:210-274``).  Neither this container nor the GPU box has
``torch_geometric``, and the hot path only ever uses attribute access, ``clone()``, ``to()`` and
``num_nodes`` on them, so this module provides a dependency-free ``Data`` with
exactly those semantics plus the PyG collation rule needed for batches of meshes.  A real PyG
``Data``/``Batch`` works wherever this one does: the models only touch attributes.
"""
from __future__ import annotations

import copy
from typing import Iterable, List

import torch


class Data:
    """Attribute bag with the subset of ``torch_geometric.data.Data`` semantics the path relies on.

    * missing attributes raise ``AttributeError`` so ``hasattr`` is False (relied on at reference
      ``src/models/Conservative.py:232``);
    * ``num_nodes`` is ``x.size(0)``, else ``pos.size(0)`` (PyG 2.6.1 ``Data.num_nodes``);
    * ``clone()`` deep-copies tensors, ``to()`` moves tensors and leaves other values alone.
    """

    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            if v is not None:
                self._store[k] = v

    # attribute protocol ---------------------------------------------------------------------
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(f"'Data' object has no attribute '{key}'")

    def __setattr__(self, key, value):
        if key == "num_nodes":
            self._store["_num_nodes"] = value
        else:
            self._store[key] = value

    def __delattr__(self, key):
        if key in self._store:
            del self._store[key]
        else:
            raise AttributeError(key)

    def __contains__(self, key):
        return key in self._store

    def keys(self):
        return [k for k in self._store.keys() if not k.startswith("_")]

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    @property
    def num_nodes(self):
        s = self._store
        if "_num_nodes" in s:
            return s["_num_nodes"]
        if "x" in s and torch.is_tensor(s["x"]):
            return s["x"].size(0)
        if "pos" in s and torch.is_tensor(s["pos"]):
            return s["pos"].size(0)
        if "face" in s and torch.is_tensor(s["face"]):
            return int(s["face"].max()) + 1
        if "edge_index" in s and torch.is_tensor(s["edge_index"]):
            return int(s["edge_index"].max()) + 1
        return None

    # copies ---------------------------------------------------------------------------------
    def clone(self):
        out = Data()
        for k, v in self._store.items():
            if torch.is_tensor(v):
                out._store[k] = v.clone()
            elif getattr(v, "__gnnfd_shared__", False):
                out._store[k] = v          # device-resident caches are shared, not copied
            else:
                out._store[k] = copy.deepcopy(v)
        return out

    def to(self, device, non_blocking: bool = False):
        out = Data()
        for k, v in self._store.items():
            out._store[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    def pin_memory(self):
        out = Data()
        for k, v in self._store.items():
            out._store[k] = v.pin_memory() if torch.is_tensor(v) else v
        return out

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = self._store[k]
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return "Data(" + ", ".join(parts) + ")"


def _is_index_key(key: str) -> bool:
    # PyG: attributes whose name contains "index" or is exactly "face" hold node ids; they are
    # concatenated along the last dim and offset by the running node count.
    return "index" in key or key == "face"


def collate(graphs: Iterable[Data], offsets: dict | None = None) -> Data:
    """Concatenate independent graphs the way PyG ``Batch.from_data_list`` does.

    ``offsets`` maps an index-valued key to the list of per-graph increments; by default the
    increment is each graph's own ``num_nodes`` (PyG ``__inc__``).  The reference's three graphs index
    into *different* node sets (``f_graph.face`` holds face ids, ``v_graph.face`` vertex ids,
    ``c_graph.edge_index`` cell ids) and PyG increments each by the *owning* graph's ``num_nodes``;
    ``collate_triplet`` below passes the counts that make the batch self-consistent.
    """
    graphs = list(graphs)
    out = Data()
    keys = graphs[0].keys()
    n_nodes = [g.num_nodes for g in graphs]
    for k in keys:
        vals = [g._store[k] for g in graphs]
        v0 = vals[0]
        if torch.is_tensor(v0):
            if _is_index_key(k):
                inc = offsets[k] if offsets and k in offsets else n_nodes
                acc, shifted = 0, []
                for v, n in zip(vals, inc):
                    shifted.append(v + acc)
                    acc += n
                out._store[k] = torch.cat(shifted, dim=-1)
            elif v0.dim() == 0:
                out._store[k] = torch.stack(vals)
            else:
                out._store[k] = torch.cat(vals, dim=0)
        else:
            out._store[k] = v0
    dev = next((v.device for v in graphs[0]._store.values() if torch.is_tensor(v)), None)   # device graphs collate on the device
    batch = torch.cat([torch.full((n,), i, dtype=torch.long, device=dev) for i, n in enumerate(n_nodes)])
    out._store["batch"] = batch
    return out


def collate_triplet(samples: List[List[Data]]) -> List[Data]:
    """Batch a list of ``[c_graph, f_graph, v_graph]`` samples element-wise (the reference's
    ``DataLoader`` collates the list of 3 graphs position by position)."""
    cs = [s[0] for s in samples]
    fs = [s[1] for s in samples]
    vs = [s[2] for s in samples]
    n_cells = [c.num_nodes for c in cs]
    n_faces = [f.num_nodes for f in fs]
    n_verts = [v.num_nodes for v in vs]
    c = collate(cs, {"edge_index": n_cells})
    f = collate(fs, {"face": n_faces})
    v = collate(vs, {"edge_index": n_verts, "face": n_verts})
    return [c, f, v]
