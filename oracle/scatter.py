"""Oracle: scatter_add and the receiver-sorted CSR (integer work in numpy).

``torch_scatter.scatter_add(src, index, dim=0, dim_size=R)`` (pinned 2.1.2 by the reference's
requirements.txt:13; call sites Mgn.py:256, Fvgn.py:314, Conservative.py:249, VertPot.py:221, ...):
``out = zeros(R, F); for i ascending: out[index[i]] += src[i]``.
"""
from __future__ import annotations

import numpy as np
import torch


def scatter_add(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """Differentiable torch-CPU form (index_add_ visits sources in ascending position on CPU)."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


def scatter_add_loop(src: np.ndarray, index: np.ndarray, dim_size: int) -> np.ndarray:
    """Pure-Python loop restatement; small cases only (cross-checks ``scatter_add``)."""
    out = np.zeros((dim_size,) + src.shape[1:], dtype=src.dtype)
    for i in range(src.shape[0]):
        out[int(index[i])] += src[i]
    return out


def csr_build(index: np.ndarray, n_rows: int):
    """Receiver-sorted CSR of an index vector.

    ``perm = argsort(index, stable)`` so a sequential per-row sum over
    ``perm[offsets[r]:offsets[r+1]]`` visits contributions in ascending source position, i.e. in the
    order the CPU ``scatter_add`` adds them.  Returns (offsets[int32, n_rows+1], perm[int32, len]).
    """
    index = np.asarray(index)
    perm = np.argsort(index, kind="stable").astype(np.int32)
    counts = np.bincount(index, minlength=n_rows).astype(np.int64)
    offsets = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return offsets.astype(np.int32), perm


def vertex_index_vector(v_edge_index: np.ndarray) -> np.ndarray:
    """cat[v_graph.edge_index[0]; v_graph.edge_index[1]]  (Fvgn.py:307-308, Mgn.py:249-250)."""
    return np.concatenate([v_edge_index[0], v_edge_index[1]])


def cell_index_vector(c_edge_index: np.ndarray) -> np.ndarray:
    """cat[col; row] = cat[edge_index[1]; edge_index[0]]  (Conservative.py:244-245)."""
    return np.concatenate([c_edge_index[1], c_edge_index[0]])
