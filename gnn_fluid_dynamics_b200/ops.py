"""Host-side wrappers of the C-ABI kernels: tensor checks in Python, then one C call each.

PyTorch is plumbing here (device memory, the current stream); all arithmetic happens in
``lib/libgnnfd_b200.so``.  Non-zero status -> ``RuntimeError`` naming the kernel.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_SILU, ACT_TANH, PRECISIONS, SEG_DIFF2, SEG_DIRECT, SEG_GATHER, SEG_MEAN3,  # noqa: F401
                   SEG_SUM2, MlpArgs, check, lib)


LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def index_narrow(index: torch.Tensor, limit: int) -> torch.Tensor:
    """int64 -> int32 with a device-side range check against [0, limit)."""
    index = _req(index.contiguous(), torch.int64, "index")
    out = torch.empty(index.shape, dtype=torch.int32, device=index.device)
    flag = torch.zeros(1, dtype=torch.int32, device=index.device)
    check(lib.gnnfd_index_narrow(index.data_ptr(), out.data_ptr(), index.numel(), int(limit),
                                 flag.data_ptr(), _stream()), "gnnfd_index_narrow")
    _count(1)
    out._gnnfd_range_flag = flag  # checked lazily by MeshTopology.validate() (needs a sync)
    return out


def csr_build(index: torch.Tensor, n_rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(offsets[n_rows+1], perm[n]) int32; perm == stable argsort(index)."""
    index = _req(index, torch.int32, "index")
    n = index.numel()
    dev = index.device
    offsets = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = lib.gnnfd_csr_workspace_bytes(n, n_rows)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    check(lib.gnnfd_csr_build(index.data_ptr(), n, n_rows, offsets.data_ptr(), perm.data_ptr(),
                              ws.data_ptr(), ws_bytes, _stream()), "gnnfd_csr_build")
    _count(6)
    return offsets, perm


def segment_sum(a: torch.Tensor, b: torch.Tensor, col_a: int, col_b: int, width: int, sign_b: float,
                offsets: torch.Tensor, perm: torch.Tensor, n_rows: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = sum over the CSR row r of a[p, col_a:+width] (p < n_half) or
    sign_b * b[p - n_half, col_b:+width], in ascending p.  n_half = a.shape[0]."""
    a = _req(a, torch.float32, "a")
    b = _req(b, torch.float32, "b")
    _req(offsets, torch.int32, "offsets")
    _req(perm, torch.int32, "perm")
    if out is None:
        out = torch.empty(n_rows, width, dtype=torch.float32, device=a.device)
    else:
        _req(out, torch.float32, "out")
    check(lib.gnnfd_segment_sum(a.data_ptr(), b.data_ptr(), a.stride(0), b.stride(0), col_a, col_b,
                                width, float(sign_b), a.shape[0], offsets.data_ptr(), perm.data_ptr(),
                                n_rows, out.data_ptr(), out.stride(0), _stream()), "gnnfd_segment_sum")
    _count(1)
    return out


@dataclass
class Seg:
    """One K-segment of an MLP input (see gnnfd_segment in include/gnnfd_b200.h)."""
    src: torch.Tensor
    mode: int = SEG_DIRECT
    idx: Sequence[torch.Tensor] = ()
    col: int = 0
    width: Optional[int] = None


@dataclass
class MLPWeights:
    """fp32 parameters of one 3-Linear MLP in PyTorch layout (+ optional tensor-core pack)."""
    w1: torch.Tensor
    b1: Optional[torch.Tensor]
    w2: torch.Tensor
    b2: Optional[torch.Tensor]
    w3: torch.Tensor
    b3: Optional[torch.Tensor]
    ln_w: Optional[torch.Tensor] = None
    ln_b: Optional[torch.Tensor] = None
    has_ln: bool = False
    ln_eps: float = 1e-5
    act: int = ACT_SILU
    packed: Optional[torch.Tensor] = None
    packed_prec: int = -1


def _fill_args(args: MlpArgs, segs: Sequence[Seg], w: MLPWeights, rows: int, precision: int):
    keep = []
    args.rows = rows
    args.n_seg = len(segs)
    k = 0
    for i, s in enumerate(segs):
        src = _req(s.src, torch.float32, f"seg[{i}].src")
        width = s.width if s.width is not None else src.shape[1] - s.col
        sg = args.seg[i]
        sg.src = src.data_ptr()
        for j in range(3):
            if j < len(s.idx):
                sg.idx[j] = _req(s.idx[j], torch.int32, f"seg[{i}].idx[{j}]").data_ptr()
            else:
                sg.idx[j] = None
        sg.ld, sg.col, sg.width, sg.mode = src.stride(0), s.col, width, s.mode
        k += width
        keep.append(src)
    args.k_in, args.hidden, args.n_out = k, w.w2.shape[0], w.w3.shape[0]
    if w.w1.shape[1] != k:
        raise RuntimeError(f"MLP input width {k} != W1.shape[1] {w.w1.shape[1]}")
    for name in ("w1", "b1", "w2", "b2", "w3", "b3", "ln_w", "ln_b"):
        t = getattr(w, name)
        setattr(args, name, _ptr(_req(t, torch.float32, name)) if t is not None else None)
    args.has_ln, args.ln_eps, args.act = int(w.has_ln), w.ln_eps, w.act
    args.precision = precision
    return keep


def pack_mlp(w: MLPWeights, precision: int) -> None:
    """(Re)build the tensor-core operand pack of ``w`` for ``precision`` in place."""
    if precision == _lib.PREC_F32:
        w.packed, w.packed_prec = None, precision
        return
    nbytes = lib.gnnfd_pack_mlp_bytes(w.w1.shape[1], w.w2.shape[0], w.w3.shape[0], precision)
    if nbytes == 0:
        raise RuntimeError("gnnfd_pack_mlp_bytes: unsupported shape/precision")
    if w.packed is None or w.packed.numel() != nbytes:
        w.packed = torch.empty(nbytes, dtype=torch.uint8, device=w.w1.device)
    args = MlpArgs()
    dummy = Seg(src=w.w1, width=w.w1.shape[1])  # geometry only; pack reads the weights
    _fill_args(args, [dummy], w, 0, precision)
    check(lib.gnnfd_pack_mlp(C.byref(args), w.packed.data_ptr(), _stream()), "gnnfd_pack_mlp")
    w.packed_prec = precision


def mlp_forward(segs: Sequence[Seg], w: MLPWeights, rows: int, precision: int = _lib.PREC_F32,
                mul: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                want_raw: bool = True, want_sum: bool = False,
                out_raw: Optional[torch.Tensor] = None, out_sum: Optional[torch.Tensor] = None):
    """Run the fused block; returns (out_raw or None, out_sum or None)."""
    args = MlpArgs()
    keep = _fill_args(args, segs, w, rows, precision)
    n_out = w.w3.shape[0]
    dev = w.w1.device
    if precision != _lib.PREC_F32:
        if w.packed is None or w.packed_prec != precision:
            pack_mlp(w, precision)
        args.packed = w.packed.data_ptr()
    if want_raw and out_raw is None:
        out_raw = torch.empty(rows, n_out, dtype=torch.float32, device=dev)
    if want_sum:
        if residual is None:
            raise RuntimeError("want_sum requires residual")
        if out_sum is None:
            out_sum = torch.empty(rows, n_out, dtype=torch.float32, device=dev)
    args.mul = _ptr(_req(mul, torch.float32, "mul")) if mul is not None else None
    args.residual = _ptr(_req(residual, torch.float32, "residual")) if residual is not None else None
    args.out_raw = _ptr(out_raw) if want_raw else None
    args.out_sum = _ptr(out_sum) if want_sum else None
    check(lib.gnnfd_mlp_forward(C.byref(args), _stream()), "gnnfd_mlp_forward")
    _count(1)
    del keep
    return (out_raw if want_raw else None), (out_sum if want_sum else None)
