"""Device-resident graph cache and on-device batch collation.

The reference builds every sample's graph triplet on the host, collates the batch with PyG and uploads ALL of it on
every step (``src/DataSet.py:210-274``, ``src/train.py:184``), although almost everything in it is static per mesh:
positions, normals, volumes, areas, the vertex-side connectivity and the ``face`` maps never change between the
timesteps of one simulation (the reference itself LRU-caches that geometry on the host, ``DataSet.py:161-172``).

Here the static part of every mesh lives in HBM, keyed by a mesh id:

* first sight of a mesh id: its static attributes are uploaded once;
* a batch (ordered tuple of mesh ids) is collated ON THE DEVICE from the resident per-mesh tensors (concatenation +
  index offsets exactly like PyG's ``Batch.from_data_list`` / ``graph.collate_triplet``; recurring batch compositions are
  memoised), so the host never concatenates anything;
* per step only the per-sample ("dynamic") attributes travel: ``c_graph.x / y``, ``f_graph.x / y`` (``x_symm`` /
  ``x_asym`` for the Conservative features) and - train mode re-flips the owner / neighbour orientation per sample
  (``src/utils/transforms.py:3-7``) - ``c_graph.edge_index`` with the face normals / fluxes that change sign with it.  They are copied from pinned host memory straight into
  their row ranges of the batch tensors (stream-ordered ``copy_(non_blocking=True)``, no host synchronisation).

The model's ``forward`` finds the batch's ``MeshTopology`` attached to the returned graphs from the previous step and
refreshes only what a new orientation invalidates (``MeshTopology.refresh_orientation``: row / col and the cell CSRs);
the vertex CSR, the 3-vertex maps and their transposed CSR are built once per batch composition.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Hashable, List, Sequence

import torch

from .graph import Data, collate_triplet

# per-sample attributes of (c_graph, f_graph, v_graph); everything else is static per mesh
# (f_graph.normal / flux change sign with the per-sample orientation flip: Conservative.py:77-79, Flux.py:71-74)
DYNAMIC = (("x", "y", "edge_index"), ("x", "y", "x_symm", "x_asym", "normal", "flux"), ())


def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


class _Batch:
    """One batch composition: up to two buffer sets (the static tensors are shared, the per-sample ones are not), so the
    next step's inputs can be uploaded on a copy stream while the current step still reads the other set."""

    def __init__(self, graphs, row_ranges, ei_offset):
        self.sets, self.row_ranges, self.ei_offset = [graphs, None], row_ranges, ei_offset
        self.last_slot = 0           # the set handed out by the latest fetch
        self.pending = None          # (slot, event) of a prefetch not yet consumed


class GraphCache:
    """``fetch(keys, host_samples)`` -> device-resident, collated ``[c_graph, f_graph, v_graph]`` of the batch.

    ``host_samples[i]`` is sample i's triplet on the host (pinned memory for asynchronous copies); ``keys[i]`` identifies
    its MESH (e.g. the simulation / geometry id of ``DataSet.py:161``): samples with the same key must share every
    attribute not listed in ``dynamic``."""

    def __init__(self, device, mesh_capacity: int = 64, batch_capacity: int = 8, dynamic: Sequence[Sequence[str]] = DYNAMIC):
        self.device = torch.device(device)      # pure data plumbing (copies + concatenation): any torch device works; the
        # models that consume the batches run on CUDA only
        self.dynamic = tuple(tuple(d) for d in dynamic)
        self.mesh_capacity, self.batch_capacity = int(mesh_capacity), int(batch_capacity)
        self._meshes: "OrderedDict[Hashable, List[Data]]" = OrderedDict()
        self._batches: "OrderedDict[tuple, _Batch]" = OrderedDict()
        self.mesh_hits = self.mesh_misses = self.batch_hits = self.batch_misses = 0
        self._copy_stream = None
        self.h2d_bytes = 0                      # bytes uploaded by the last fetch
        self.h2d_static_bytes = 0               # ... of which static attributes of meshes seen for the first time

    # ------------------------------------------------------------------------------------------------ meshes
    def _mesh(self, key, sample) -> List[Data]:
        m = self._meshes.get(key)
        if m is not None:
            self.mesh_hits += 1
            self._meshes.move_to_end(key)
            return m
        self.mesh_misses += 1
        m = []
        for gi, g in enumerate(sample):
            d = Data()
            for k, v in g._store.items():
                if torch.is_tensor(v):
                    if k in self.dynamic[gi] or k == "batch":
                        continue
                    d._store[k] = v.to(self.device, non_blocking=True)
                    self.h2d_static_bytes += _nbytes(v)
                else:
                    d._store[k] = v
            # node counts the collation needs even when the attribute that defines them is dynamic
            d.num_nodes = g.num_nodes
            m.append(d)
        self._meshes[key] = m
        while len(self._meshes) > self.mesh_capacity:
            self._meshes.popitem(last=False)
        return m

    # ----------------------------------------------------------------------------------------------- batches
    def _assemble(self, keys, host_samples) -> _Batch:
        meshes = [self._mesh(k, s) for k, s in zip(keys, host_samples)]
        graphs = collate_triplet(meshes)                       # device-side concatenation of the static attributes
        counts = [[g.num_nodes for g in s] for s in host_samples]
        row_ranges, ei_offset = [], None
        for gi in range(3):
            offs, acc = [], 0
            for c in counts:
                offs.append((acc, acc + c[gi]))
                acc += c[gi]
            row_ranges.append(offs)
            for name in self.dynamic[gi]:
                src = host_samples[0][gi]._store.get(name)
                if src is None or not torch.is_tensor(src):
                    continue
                if name == "edge_index":                       # [2, E_i] cell ids, concatenated along the last dim
                    e_counts = [s[gi]._store[name].shape[1] for s in host_samples]
                    graphs[gi]._store[name] = torch.empty(2, sum(e_counts), dtype=src.dtype, device=self.device)
                    ei_offset = torch.cat([torch.full((e,), r0, dtype=src.dtype, device=self.device)
                                           for e, (r0, _) in zip(e_counts, offs)])
                else:
                    graphs[gi]._store[name] = torch.empty((acc,) + tuple(src.shape[1:]), dtype=src.dtype, device=self.device)
        for g in graphs:
            g._store.pop("_num_nodes", None)                  # the batch's counts follow from its tensors again
        return _Batch(graphs, row_ranges, ei_offset)

    def _batch(self, keys, host_samples) -> _Batch:
        b = self._batches.get(keys)
        if b is None:
            self.batch_misses += 1
            b = self._assemble(keys, host_samples)
            self._batches[keys] = b
            while len(self._batches) > self.batch_capacity:
                self._batches.popitem(last=False)
        else:
            self.batch_hits += 1
            self._batches.move_to_end(keys)
        return b

    def fetch(self, keys: Sequence[Hashable], host_samples: Sequence[Sequence[Data]]) -> List[Data]:
        keys = tuple(keys)
        if len(keys) != len(host_samples):
            raise ValueError("one mesh key per sample")
        self.h2d_static_bytes = 0
        b = self._batch(keys, host_samples)
        token = tuple(id(g) for s in host_samples for g in s)
        if b.pending is not None and b.pending[3] != token:
            # a prefetch of OTHER samples is in flight for this batch composition: let it land, then upload these
            torch.cuda.current_stream(self.device).wait_event(b.pending[1])
            b.pending = None
        if b.pending is not None:                  # uploaded ahead of time by prefetch(): just order this stream after it
            slot, event, moved, _ = b.pending
            b.pending = None
            torch.cuda.current_stream(self.device).wait_event(event)
        else:
            slot = b.last_slot
            moved = self._upload(b, slot, host_samples)
        b.last_slot = slot
        self.h2d_bytes = moved + self.h2d_static_bytes
        return b.sets[slot]

    def prefetch(self, keys: Sequence[Hashable], host_samples: Sequence[Sequence[Data]], after=None) -> None:
        """Start uploading the per-sample attributes of the NEXT batch on the cache's copy stream, into the buffer set the
        model is not reading; the matching ``fetch`` then only waits for the copy (a loader's double buffering: the
        host->device traffic of step k + 1 overlaps the compute of step k).  ``after``: a CUDA event recorded on the compute
        stream BEFORE the current step's kernels were enqueued - the copies then wait only for the work before it (the step
        that last read the target set), so a caller can enqueue the current step first and the prefetch afterwards, keeping
        the ~50 small copy calls off the host's critical path."""
        if self.device.type != "cuda":
            return
        keys = tuple(keys)
        b = self._batch(keys, host_samples)
        slot = 1 - b.last_slot
        if b.sets[slot] is None:                   # second buffer set: shares every static tensor with the first
            twin = []
            for gi, g in enumerate(b.sets[b.last_slot]):
                d = Data()
                for k, v in g._store.items():
                    if k == "topology":
                        continue
                    d._store[k] = torch.empty_like(v) if (torch.is_tensor(v) and k in self.dynamic[gi]) else v
                twin.append(d)
            b.sets[slot] = twin
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        free = after
        if free is None:
            free = torch.cuda.Event()
            free.record(torch.cuda.current_stream(self.device))     # everything enqueued so far (the step that last read this set)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(free)
            moved = self._upload(b, slot, host_samples)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        b.pending = (slot, done, moved, tuple(id(g) for s in host_samples for g in s))

    def _upload(self, b: _Batch, slot: int, host_samples) -> int:
        """Per-sample attributes of every sample -> their row ranges of buffer set ``slot`` (on the current stream)."""
        graphs = b.sets[slot]
        moved = 0
        for gi in range(3):
            for name in self.dynamic[gi]:
                dst = graphs[gi]._store.get(name)
                if dst is None:
                    continue
                if name == "edge_index":
                    e0 = 0
                    for s in host_samples:
                        src = s[gi]._store[name]
                        dst[:, e0:e0 + src.shape[1]].copy_(src, non_blocking=True)
                        e0 += src.shape[1]
                        moved += _nbytes(src)
                    if e0 != dst.shape[1]:
                        raise RuntimeError("GraphCache: face count changed under the same mesh keys")
                    dst.add_(b.ei_offset)                      # PyG's per-graph increment, on the device
                else:
                    for s, (r0, r1) in zip(host_samples, b.row_ranges[gi]):
                        src = s[gi]._store[name]
                        if src.shape[0] != r1 - r0:
                            raise RuntimeError(f"GraphCache: graph {gi}.{name} changed its row count under the same mesh key")
                        dst[r0:r1].copy_(src, non_blocking=True)
                        moved += _nbytes(src)
        return moved

    def clear(self):
        self._meshes.clear()
        self._batches.clear()
