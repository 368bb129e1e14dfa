"""Host-side wrappers of the C-ABI kernels: tensor checks in Python, then one C call each.

PyTorch is plumbing here (device memory, the current stream); all arithmetic happens in
``lib/libgnnfd_b200.so``.  Non-zero status -> ``RuntimeError`` naming the kernel.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_SILU, ACT_TANH, PRECISIONS, SEG_DIFF2, SEG_DIRECT, SEG_GATHER, SEG_MEAN3,  # noqa: F401
                   SEG_SUM2, MlpArgs, MlpBackwardArgs, WgradArgs, check, lib)


LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


# A training step issues ~250 library calls and ~1 800 argument checks; torch.cuda.current_stream() builds a Stream object
# and torch.cuda.current_device() walks the lazy-init guard on every call (2 + 2 ms of the ~15 ms a step takes to enqueue,
# scripts/prof_host.py).  The raw getters below are what those wrappers end in.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _current_device() -> int:
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if t.device.index != _current_device():
        # the C launches go to the CURRENT device's stream: a tensor on another GPU would fault or be reached through
        # peer access silently (use torch.cuda.set_device / `with torch.cuda.device(...)` around the model call)
        raise RuntimeError(f"{name}: tensor is on cuda:{t.device.index} but the current device is "
                           f"cuda:{torch.cuda.current_device()}")
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
    elif not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.stride(1) == 1):
        raise RuntimeError("segment_sum: out must be a CUDA fp32 matrix with unit column stride (column slices are fine)")
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
    split: Optional[torch.Tensor] = None   # 16-bit hi|lo shadow of src ([rows, 2 * ld]) written by mlp_forward(out_split=...)


def split_dtype(precision: int):
    """Element type of the split shadow for a tensor-core precision (None: the precision has no split operands)."""
    if precision == _lib.PREC_BF16X3:
        return torch.bfloat16
    if precision == _lib.PREC_FP16X3:
        return torch.float16
    return None


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
    bwd_packs: Optional[dict] = None   # dgrad operand packs keyed by (which, col0, precision)
    # training-mode dropout of the two hidden activations (gnnfd_mlp_args.dropout_p): w2 / w3 above are then the module's
    # weights ALREADY divided by (1 - drop_p) and mlp_backward scales their gradients by the same factor
    drop_p: float = 0.0


@dataclass
class MLPStash:
    """What the backward needs from one fused-MLP forward (the autograd stash): pre-activations of the
    two hidden layers, the normalised rows before the LayerNorm affine and 1/sqrt(var + eps)."""
    a1: torch.Tensor
    a2: torch.Tensor
    xhat: Optional[torch.Tensor]
    rstd: Optional[torch.Tensor]


STATIC_OPERANDS = os.environ.get("GNNFD_STATIC_OPERANDS", "1") != "0"      # A/B knob of the early operand fetch


def dropout_seed() -> int:
    """A fresh 63-bit seed from torch's default CPU generator (so ``torch.manual_seed`` makes training runs repeatable);
    no device synchronisation."""
    return int(torch.empty((), dtype=torch.int64).random_().item()) & (2 ** 63 - 1)


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
        sg.src_rows = src.shape[0]
        sg.split = None
        if s.split is not None and split_dtype(precision) is not None:
            sp = s.split
            if not (sp.is_cuda and sp.dtype == split_dtype(precision) and sp.is_contiguous()
                    and tuple(sp.shape) == (src.shape[0], 2 * src.stride(0))):
                raise RuntimeError(f"seg[{i}].split: expected a contiguous {split_dtype(precision)} [{src.shape[0]}, "
                                   f"{2 * src.stride(0)}] CUDA tensor, got {sp.dtype} {tuple(sp.shape)}")
            sg.split = sp.data_ptr()
            keep.append(sp)
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
                out_raw: Optional[torch.Tensor] = None, out_sum: Optional[torch.Tensor] = None,
                stash: bool = False, peer: Optional[tuple] = None,
                out_split: Optional[torch.Tensor] = None, split_of_sum: bool = False):
    """Run the fused block; returns (out_raw or None, out_sum or None) and, with ``stash=True``
    (training), additionally the ``MLPStash`` the backward consumes.

    ``out_sum`` may be ``residual`` itself: the residual stream is then updated in place (the fast inference path - the
    add is done by the TMA store).  ``out_split`` [rows, 256] (``split_dtype(precision)``) receives the 16-bit hi|lo
    shadow of the raw output (or of the sum with ``split_of_sum``) that the next block's gathers consume through
    ``Seg.split``."""
    args = MlpArgs()
    keep = _fill_args(args, segs, w, rows, precision)
    n_out = w.w3.shape[0]
    dev = w.w1.device
    if precision != _lib.PREC_F32:
        packed_now = w.packed is None or w.packed_prec != precision
        if packed_now:
            pack_mlp(w, precision)
        args.packed = w.packed.data_ptr()
        # Inside a CUDA-graph capture, operands produced BEFORE the capture (the pack of an earlier call, the parameters,
        # the topology's index arrays) cannot be written by anything in flight when the replayed launch starts: the kernel
        # may fetch them ahead of griddepcontrol.wait (gnnfd_mlp_args.static_operands; inference launches only)
        args.static_operands = int(STATIC_OPERANDS and not packed_now and not stash and not torch.is_grad_enabled()
                                   and torch.cuda.is_current_stream_capturing())
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
    if out_split is not None:
        if not (out_split.is_cuda and out_split.dtype == split_dtype(precision) and out_split.is_contiguous()
                and tuple(out_split.shape) == (rows, 2 * n_out)):
            raise RuntimeError(f"out_split: expected a contiguous {split_dtype(precision)} [{rows}, {2 * n_out}] CUDA tensor")
        args.out_split, args.split_of_sum = out_split.data_ptr(), int(split_of_sum)
    if peer is not None:          # (peer matrices, shift): GATHER indices are (peer << shift) | row
        bases, shift = peer
        for i, t in enumerate(bases):
            args.peer_base[i] = t.data_ptr() if t is not None else None
        args.peer_shift = shift
    st = None
    if stash:
        if precision == _lib.PREC_F32:
            raise RuntimeError("training needs a tensor-core precision (the f32 kernel keeps no stash)")
        hid = w.w2.shape[0]
        st = MLPStash(a1=torch.empty(rows, hid, dtype=torch.float32, device=dev),
                      a2=torch.empty(rows, hid, dtype=torch.float32, device=dev),
                      xhat=torch.empty(rows, n_out, dtype=torch.float32, device=dev) if w.has_ln else None,
                      rstd=torch.empty(rows, dtype=torch.float32, device=dev) if w.has_ln else None)
        args.save_a1, args.save_a2 = st.a1.data_ptr(), st.a2.data_ptr()
        args.save_xhat, args.save_rstd = _ptr(st.xhat), _ptr(st.rstd)
    if w.drop_p > 0.0:
        args.dropout_p, args.dropout_seed = w.drop_p, dropout_seed()
    check(lib.gnnfd_mlp_forward(C.byref(args), _stream()), "gnnfd_mlp_forward")
    _count(1)
    del keep
    if stash:
        return (out_raw if want_raw else None), (out_sum if want_sum else None), st
    return (out_raw if want_raw else None), (out_sum if want_sum else None)


# ------------------------------------------------------------------------------------------ backward

def _fill_segment(sg, s: Seg, name: str):
    src = _req(s.src, torch.float32, f"{name}.src")
    width = s.width if s.width is not None else src.shape[1] - s.col
    sg.src = src.data_ptr()
    for j in range(3):
        sg.idx[j] = _req(s.idx[j], torch.int32, f"{name}.idx[{j}]").data_ptr() if j < len(s.idx) else None
    sg.ld, sg.col, sg.width, sg.mode = src.stride(0), s.col, width, s.mode
    return width


def linear_tc(src: Seg, rows: int, w_t: torch.Tensor, ld_n: int, ld_k: int, w_rows: int, k_in: int,
              pack_cache: dict, pack_key, precision: int, bias: Optional[torch.Tensor] = None,
              mul: Optional[torch.Tensor] = None, mul_mode: int = 0, residual: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[rows,128] = (src . Wt^T (+ bias)) (* mul | act'(mul)) (+ residual), one tensor-core Linear
    (gnnfd_mlp_forward with n_layers = 1).  ``w_t`` is addressed as element (n, k) = w_t.data_ptr()
    [n * ld_n + k * ld_k] so a forward weight is used transposed in place (dgrad: dH = dA . W)."""
    args = MlpArgs()
    args.rows, args.n_seg = rows, 1
    width = _fill_segment(args.seg[0], src, "src")
    if width != k_in:
        raise RuntimeError(f"linear_tc: source width {width} != k_in {k_in}")
    args.k_in, args.hidden, args.n_out = k_in, 128, 128
    args.w1 = w_t.data_ptr()
    args.b1 = _ptr(bias)
    args.has_ln, args.ln_eps, args.act = 0, 0.0, ACT_SILU
    args.precision = precision
    args.n_layers, args.mul_mode = 1, mul_mode
    args.w1_ld_n, args.w1_ld_k, args.w1_rows = ld_n, ld_k, w_rows
    if out is None:
        out = torch.empty(rows, 128, dtype=torch.float32, device=w_t.device)
    args.mul = _ptr(_req(mul, torch.float32, "mul")) if mul is not None else None
    if residual is not None:
        args.residual, args.out_sum = _req(residual, torch.float32, "residual").data_ptr(), out.data_ptr()
    else:
        args.out_raw = out.data_ptr()
    key = (pack_key, precision)
    packed = pack_cache.get(key)
    if packed is None:
        nbytes = lib.gnnfd_pack_mlp_bytes(k_in, 128, 128, precision)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=w_t.device)
        check(lib.gnnfd_pack_mlp(C.byref(args), packed.data_ptr(), _stream()), "gnnfd_pack_mlp")
        _count(1)
        pack_cache[key] = packed
    args.packed = packed.data_ptr()
    check(lib.gnnfd_mlp_forward(C.byref(args), _stream()), "gnnfd_mlp_forward")
    _count(1)
    return out


def mlp_backward_workspace(rows: int, device) -> torch.Tensor:
    """Scratch for ``mlp_backward`` calls of up to ``rows`` rows (three [rows,128] matrices + split-K partials)."""
    args = MlpArgs()
    args.rows, args.n_seg, args.has_ln, args.precision, args.n_out = rows, 3, 1, _lib.PREC_BF16X3, 128
    return torch.empty(lib.gnnfd_mlp_backward_workspace_bytes(C.byref(args)), dtype=torch.uint8, device=device)


def mlp_backward(segs: Sequence[Seg], w: MLPWeights, st: "MLPStash", rows: int, g: torch.Tensor, precision: int,
                 din: Sequence[Optional[dict]], workspace: torch.Tensor, da1_out: Optional[torch.Tensor] = None,
                 skip_wgrad_l1: int = 0):
    """Whole backward of one fused MLP in one C call (gnnfd_mlp_backward).  ``din[i]`` is None or a dict with
    optional ``residual`` / ``out``.  Returns ({name: grad} for w1,b1,w2,b2,w3,b3,ln_w,ln_b present, [dIn_i])."""
    b = MlpBackwardArgs()
    keep = _fill_args(b.fwd, segs, w, rows, precision)
    dev = w.w1.device
    chain = int(len(din) > 0 and din[0] is not None)
    if w.bwd_packs is None or w.bwd_packs.get("key") != (precision, chain):
        nbytes = lib.gnnfd_pack_mlp_backward_bytes(C.byref(b.fwd))
        pk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.gnnfd_pack_mlp_backward(C.byref(b.fwd), pk.data_ptr(), chain, _stream()), "gnnfd_pack_mlp_backward")
        _count(len(segs) if chain else 2 + len(segs))
        w.bwd_packs = {"key": (precision, chain), "pack": pk}
    b.packed_bwd = w.bwd_packs["pack"].data_ptr()
    g = _req(g.contiguous(), torch.float32, "g")
    b.g = g.data_ptr()
    b.a1, b.a2, b.xhat, b.rstd = st.a1.data_ptr(), st.a2.data_ptr(), _ptr(st.xhat), _ptr(st.rstd)
    n_out, k_in = w.w3.shape[0], w.w1.shape[1]
    # every parameter gradient of the MLP in one flat allocation
    sizes = [("w1", 128 * k_in), ("b1", 128 if w.b1 is not None else 0), ("w2", 128 * 128),
             ("b2", 128 if w.b2 is not None else 0), ("w3", n_out * 128), ("b3", n_out if w.b3 is not None else 0),
             ("ln_w", n_out if w.ln_w is not None else 0), ("ln_b", n_out if w.ln_b is not None else 0)]
    flat = torch.empty(sum(n for _, n in sizes), dtype=torch.float32, device=dev)
    grads, o = {}, 0
    shapes = {"w1": (128, k_in), "w2": (128, 128), "w3": (n_out, 128)}
    for name, n in sizes:
        if n:
            grads[name] = flat[o:o + n].view(shapes.get(name, (n,)))
            setattr(b, "d_" + name, flat.data_ptr() + 4 * o)
            o += n
    dins = []
    for i in range(len(segs)):
        spec = din[i] if i < len(din) else None
        if spec is None:
            dins.append(None)
            continue
        out = spec.get("out")
        if out is None:
            out = torch.empty(rows, 128, dtype=torch.float32, device=dev)
        b.din_out[i] = out.data_ptr()
        res = spec.get("residual")
        b.din_residual[i] = _req(res, torch.float32, "residual").data_ptr() if res is not None else None
        dins.append(out)
    b.workspace, b.workspace_bytes = workspace.data_ptr(), workspace.numel()
    b.da1_out, b.skip_wgrad_l1 = _ptr(da1_out), int(skip_wgrad_l1)     # see gnnfd_mlp_backward_args
    check(lib.gnnfd_mlp_backward(C.byref(b), _stream()), "gnnfd_mlp_backward")
    _count((2 if w.has_ln else 0) + 6 + 2 + sum(1 for d in dins if d is not None))
    del keep
    if w.drop_p > 0.0:      # w2 / w3 were the module's weights / (1 - p): chain rule of that substitution
        grads["w2"].mul_(1.0 / (1.0 - w.drop_p))
        grads["w3"].mul_(1.0 / (1.0 - w.drop_p))
    return grads, dins


def dgrad_chain(w: MLPWeights, st: "MLPStash", dy: torch.Tensor, seg0_width: int, precision: int,
                residual: Optional[torch.Tensor] = None, pack_cache: Optional[dict] = None):
    """The dgrad chain of one MLP as ONE tensor-core pass (gnnfd_mlp_forward with bwd_chain = 1):
    dA2 = (dy W3) * act'(a2), dA1 = (dA2 W2) * act'(a1), dIn0 = dA1 W1[:, 0:seg0_width] (+ residual).
    Returns (dIn0, dA2, dA1).  gnnfd_mlp_backward issues exactly this launch; exposed for tests and profiling."""
    dy = _req(dy, torch.float32, "dy")
    rows, n_out, k_in = dy.shape[0], w.w3.shape[0], w.w1.shape[1]
    dev = dy.device
    args = MlpArgs()
    args.rows, args.n_seg = rows, 1
    _fill_segment(args.seg[0], Seg(dy), "dy")
    args.k_in, args.hidden, args.n_out = n_out, 128, 128
    args.w1, args.w1_ld_n, args.w1_ld_k, args.w1_rows = w.w3.data_ptr(), 1, 128, 128
    args.w2, args.w2_ld_n, args.w2_ld_k = w.w2.data_ptr(), 1, 128
    args.w3, args.w3_ld_n, args.w3_ld_k, args.w3_rows = w.w1.data_ptr(), 1, k_in, seg0_width
    args.act, args.precision, args.n_layers, args.bwd_chain = w.act, precision, 3, 1
    args.hid_mul1, args.hid_mul2 = st.a2.data_ptr(), st.a1.data_ptr()
    d_a2 = torch.empty(rows, 128, dtype=torch.float32, device=dev)
    d_a1 = torch.empty(rows, 128, dtype=torch.float32, device=dev)
    out = torch.empty(rows, 128, dtype=torch.float32, device=dev)
    args.save_a1, args.save_a2 = d_a2.data_ptr(), d_a1.data_ptr()
    if residual is not None:
        args.residual, args.out_sum = _req(residual, torch.float32, "residual").data_ptr(), out.data_ptr()
    else:
        args.out_raw = out.data_ptr()
    cache = pack_cache if pack_cache is not None else {}
    pk = cache.get(("chain", precision))
    if pk is None:
        pk = torch.empty(lib.gnnfd_pack_mlp_bytes(n_out, 128, 128, precision), dtype=torch.uint8, device=dev)
        check(lib.gnnfd_pack_mlp(C.byref(args), pk.data_ptr(), _stream()), "gnnfd_pack_mlp")
        _count(1)
        cache[("chain", precision)] = pk
    args.packed = pk.data_ptr()
    check(lib.gnnfd_mlp_forward(C.byref(args), _stream()), "gnnfd_mlp_forward")
    _count(1)
    return out, d_a2, d_a1


def ln_backward(g: torch.Tensor, xhat: torch.Tensor, rstd: torch.Tensor, ln_w: Optional[torch.Tensor]):
    """(dy[rows,128], sums[3,128]) - see gnnfd_ln_backward."""
    g = _req(g, torch.float32, "g")
    rows = g.shape[0]
    dy = torch.empty_like(g)
    sums = torch.empty(3, 128, dtype=torch.float32, device=g.device)
    nb = lib.gnnfd_ln_backward_workspace_bytes(rows)
    ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
    check(lib.gnnfd_ln_backward(g.data_ptr(), _req(xhat, torch.float32, "xhat").data_ptr(),
                                _req(rstd, torch.float32, "rstd").data_ptr(), _ptr(ln_w), rows, dy.data_ptr(),
                                sums.data_ptr(), ws.data_ptr(), nb, _stream()), "gnnfd_ln_backward")
    _count(2)
    return dy, sums


def wgrad(a: Seg, b: Sequence[Seg], rows: int, out: torch.Tensor, a_act: int = 0, b_act: int = 0,
          transpose_out: bool = False, colsum: Optional[torch.Tensor] = None, colsum_of_b: bool = False,
          workspace: Optional[torch.Tensor] = None, single_pass: bool = False) -> torch.Tensor:
    """out[m, n] = sum_r A[r, m] B[r, n] on the tensor cores (gnnfd_wgrad); activations: 0 none, 1 SiLU, 2 tanh."""
    args = WgradArgs()
    args.rows = rows
    _fill_segment(args.a, a, "a")
    args.a_act, args.b_act, args.n_b = a_act, b_act, len(b)
    n_pad, g = 0, (32 if single_pass else 64)          # column padding of the operand images (TF32 / split-bf16 atoms)
    for i, s in enumerate(b):
        n_pad += (_fill_segment(args.b[i], s, f"b[{i}]") + g - 1) // g * g
    if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.stride(1) == 1):
        raise RuntimeError("wgrad: out must be a CUDA fp32 matrix with unit column stride (column blocks are fine)")
    args.out, args.ld_out, args.transpose_out = out.data_ptr(), out.stride(0), int(transpose_out)
    args.colsum, args.colsum_of_b = _ptr(colsum), int(colsum_of_b)
    args.precision = 1 if single_pass else 0
    nb = lib.gnnfd_wgrad_workspace_bytes(rows, n_pad)
    if workspace is None or workspace.numel() < nb:
        workspace = torch.empty(nb, dtype=torch.uint8, device=out.device)
    check(lib.gnnfd_wgrad(C.byref(args), workspace.data_ptr(), workspace.numel(), _stream()), "gnnfd_wgrad")
    _count(2)
    return out


def segment_sum3(a, b, c, cols, width: int, sign_b: float, n_part: int, offsets, perm, n_rows: int,
                 scale: float = 1.0, base: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """out[r] = base[r] + scale * sum over CSR row r of the three-part source (gnnfd_segment_sum3)."""
    a = _req(a, torch.float32, "a")
    ld = a.stride(0)
    for t in (b, c):
        if t is not None and _req(t, torch.float32, "part").stride(0) != ld:
            raise RuntimeError("segment_sum3: parts must share a row stride")
    if out is None:
        out = torch.empty(n_rows, width, dtype=torch.float32, device=a.device)
    check(lib.gnnfd_segment_sum3(a.data_ptr(), _ptr(b), _ptr(c), ld, cols[0], cols[1], cols[2], width, float(sign_b),
                                 n_part, _req(offsets, torch.int32, "offsets").data_ptr(),
                                 _req(perm, torch.int32, "perm").data_ptr(), n_rows, float(scale), _ptr(base),
                                 base.stride(0) if base is not None else 0, out.data_ptr(), out.stride(0), _stream()),
          "gnnfd_segment_sum3")
    _count(1)
    return out


def gather_pair_add(src: torch.Tensor, i0: torch.Tensor, i1: torch.Tensor, sign: float, halves: bool, rows: int,
                    base: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[rows,128] = base + gathered pair (gnnfd_gather_pair_add); ``out`` may be ``base`` (in place)."""
    src = _req(src, torch.float32, "src")
    if out is None:
        out = torch.empty(rows, 128, dtype=torch.float32, device=src.device)
    check(lib.gnnfd_gather_pair_add(out.data_ptr(), _ptr(base), src.data_ptr(), src.stride(0),
                                    _req(i0, torch.int32, "i0").data_ptr(), _req(i1, torch.int32, "i1").data_ptr(),
                                    float(sign), int(halves), rows, _stream()), "gnnfd_gather_pair_add")
    _count(1)
    return out


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = src[idx[r]] (halo pack, gnnfd_gather_rows)."""
    src = _req(src, torch.float32, "src")
    idx = _req(idx, torch.int32, "idx")
    n, width = idx.numel(), src.shape[1]
    if out is None:
        out = torch.empty(n, width, dtype=torch.float32, device=src.device)
    check(lib.gnnfd_gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), n, width, out.data_ptr(), _stream()),
          "gnnfd_gather_rows")
    _count(1)
    return out


def gather_cols_add(dst: torch.Tensor, col: int, width: int, src: torch.Tensor, idx: torch.Tensor, scale: float = 1.0):
    """dst[k, col:col+width] += scale * src[idx[k], :width]  in place (gnnfd_gather_cols_add)."""
    dst = _req(dst, torch.float32, "dst")
    src = _req(src, torch.float32, "src")
    check(lib.gnnfd_gather_cols_add(dst.data_ptr(), dst.stride(0), col, width, src.data_ptr(), src.stride(0),
                                    _req(idx, torch.int32, "idx").data_ptr(), float(scale), dst.shape[0], _stream()),
          "gnnfd_gather_cols_add")
    _count(1)
    return dst
