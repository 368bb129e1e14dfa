"""Autograd-aware wrappers of the two kernel families the processors are written with - the fused MLP block and the
deterministic segment sum - so that EVERY family's data-flow in ``processor.py`` is trainable as written.

When no gradient is being recorded (inference, rollout) the wrappers are a plain kernel call.  Under autograd each
call is one ``torch.autograd.Function`` whose backward uses the same backward kernels as the hand-scheduled training
path (``training.mlp_backward`` = one ``gnnfd_mlp_backward`` call; transposed gathers as deterministic CSR segment
sums; the segment sum's transpose as a gather), only without the cross-op fusions of ``training.EncodeProcessDecode``
(which stays the fast path for the Fvgn / Mgn / VertPot families).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from ._lib import ACT_SILU, SEG_DIFF2, SEG_DIRECT, SEG_MEAN3
from .ops import Seg

_NAMES = ("w1", "b1", "w2", "b2", "w3", "b3", "ln_w", "ln_b")


def _needs_grad(tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors)


def _transpose_csr(indices: Sequence[torch.Tensor], n_rows: int):
    """CSR of cat[indices] over ``n_rows`` rows, cached on the first index tensor object (topology tensors are
    long-lived Python objects, so the cache lives exactly as long as the index it describes)."""
    first = indices[0]
    cache = getattr(first, "_gnnfd_tcsr", None)
    if cache is None:
        cache = {}
        first._gnnfd_tcsr = cache
    key = (tuple(id(t) for t in indices[1:]), n_rows)
    if key not in cache:
        flat = indices[0] if len(indices) == 1 else torch.cat(list(indices))
        cache[key] = ops.csr_build(flat, n_rows)
    return cache[key]


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, *tensors):
        seq, act, segs_meta, rows, prec, want_raw, want_sum = meta
        from .processor import weights_of
        params, rest = tensors[:8], tensors[8:]
        n_seg = len(segs_meta)
        srcs, mul, residual = rest[:n_seg], rest[n_seg], rest[n_seg + 1]
        segs = [Seg(s.detach(), mode, idx, col, width) for s, (mode, idx, col, width) in zip(srcs, segs_meta)]
        w = weights_of(seq, act)
        raw, summed, st = ops.mlp_forward(segs, w, rows, prec, mul=None if mul is None else mul.detach(),
                                          residual=None if residual is None else residual.detach(),
                                          want_raw=want_raw, want_sum=want_sum, stash=True)
        ctx.w, ctx.st, ctx.segs, ctx.rows, ctx.prec = w, st, segs, rows, prec
        ctx.mul = None if mul is None else mul.detach()
        ctx.src_rows = [s.shape for s in srcs]
        ctx.has_res = residual is not None
        ctx.need = [t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors]
        outs = tuple(o for o in (raw, summed))
        return outs

    @staticmethod
    def backward(ctx, g_raw, g_sum):
        from .training import mlp_backward
        w, st, segs, rows, prec = ctx.w, ctx.st, ctx.segs, ctx.rows, ctx.prec
        n_seg = len(segs)
        g = g_raw if g_sum is None else (g_sum if g_raw is None else g_raw + g_sum)
        if g is None:
            return (None,) * (1 + 8 + n_seg + 2)
        g = g.contiguous()
        d_mul = None
        if ctx.mul is not None:       # out = LN(...) * mul  (ConservativeA block 0 / ConservativeI keep matrix)
            if ctx.need[8 + n_seg]:
                ln_out = st.xhat if w.ln_w is None else st.xhat * w.ln_w + w.ln_b
                d_mul = g * ln_out
            g = g * ctx.mul
        need_src = ctx.need[8:8 + n_seg]
        ws = ops.mlp_backward_workspace(rows, g.device)
        grads, dins = mlp_backward(w, st, segs, rows, g, prec, [({} if n else None) for n in need_src], ws)
        d_srcs = []
        for seg, din, shape, need in zip(segs, dins, ctx.src_rows, need_src):
            if not need:
                d_srcs.append(None)
                continue
            width = seg.width if seg.width is not None else shape[1] - seg.col
            n_src = shape[0]
            if seg.mode == SEG_DIRECT:
                if seg.col == 0 and width == shape[1] and width == din.shape[1]:
                    d = din
                else:
                    d = torch.zeros(shape, dtype=torch.float32, device=din.device)
                    d[:, seg.col:seg.col + width] = din[:, :width]
            else:
                # transpose of the gather: deterministic segment sum of the row gradients over the CSR of the indices
                idx = list(seg.idx)
                off, perm = _transpose_csr(idx, n_src)
                sign = -1.0 if seg.mode == SEG_DIFF2 else 1.0
                scale = 1.0 / 3.0 if seg.mode == SEG_MEAN3 else 1.0
                part = ops.segment_sum3(din, din, din, (0, 0, 0), width, sign,
                                        rows, off, perm, n_src, scale=scale)
                if seg.col == 0 and part.shape[1] == shape[1]:
                    d = part
                else:
                    d = torch.zeros(shape, dtype=torch.float32, device=din.device)
                    d[:, seg.col:seg.col + width] = part[:, :width]
            d_srcs.append(d)
        d_res = g_sum if (ctx.has_res and ctx.need[8 + n_seg + 1]) else None
        p_grads = [gr if need else None for gr, need in zip(grads, ctx.need[:8])]
        ctx.st = None
        return (None, *p_grads, *d_srcs, d_mul, d_res)


def mlp(seq, segs: Sequence[Seg], rows: int, prec: int, act: int = ACT_SILU, mul: Optional[torch.Tensor] = None,
        residual: Optional[torch.Tensor] = None, want_raw: bool = True, want_sum: bool = False,
        inplace: bool = False, out_split: Optional[torch.Tensor] = None, split_of_sum: bool = False):
    """Fused MLP block of module ``seq`` -> (out_raw or None, out_sum or None); differentiable w.r.t. the module's
    parameters, the segment sources, ``mul`` and ``residual`` when autograd is recording.

    Inference only (``processor.inference_mode``): ``inplace`` writes the sum into ``residual`` itself, ``out_split``
    receives the 16-bit split shadow of the output for the next block's TMA gathers (see ``ops.mlp_forward``)."""
    from .processor import weights_of
    from .training import Site
    site_params = Site(seq, act).params
    srcs = [s.src for s in segs]
    if not _needs_grad(list(site_params) + srcs + [mul, residual]):
        return ops.mlp_forward(segs, weights_of(seq, act), rows, prec, mul=mul, residual=residual,
                               want_raw=want_raw, want_sum=want_sum,
                               out_sum=residual if (inplace and want_sum) else None,
                               out_split=out_split, split_of_sum=split_of_sum)
    if inplace or out_split is not None or any(s.split is not None for s in segs):
        raise RuntimeError("in-place residuals / split shadows are inference-only (autograd is recording)")
    meta = (seq, act, [(s.mode, tuple(s.idx), s.col, s.width) for s in segs], rows, prec, want_raw, want_sum)
    raw, summed = _MLPFn.apply(meta, *site_params, *srcs, mul, residual)
    return raw, summed


class _SegSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, meta):
        col_a, col_b, width, sign_b, offsets, perm, n_rows, i_first, i_second = meta
        s = src.detach()
        out = ops.segment_sum(s, s, col_a, col_b, width, sign_b, offsets, perm, n_rows)
        ctx.meta, ctx.shape = meta, src.shape
        return out

    @staticmethod
    def backward(ctx, g):
        col_a, col_b, width, sign_b, offsets, perm, n_rows, i_first, i_second = ctx.meta
        g = g.contiguous()
        d = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        ops.gather_cols_add(d, col_a, width, g, i_first, 1.0)          # position p < E contributed src[p, col_a:]
        ops.gather_cols_add(d, col_b, width, g, i_second, sign_b)      # position p >= E contributed sign * src[p-E, col_b:]
        return d, None


def segment_sum(src: torch.Tensor, col_a: int, col_b: int, width: int, sign_b: float, offsets, perm, n_rows: int,
                i_first: torch.Tensor, i_second: torch.Tensor):
    """out[r] = sum_{i_first[k] = r} src[k, col_a:+w] + sign_b * sum_{i_second[k] = r} src[k, col_b:+w]  over the
    receiver-sorted CSR (offsets, perm) of cat[i_first; i_second]; differentiable w.r.t. ``src``."""
    if not _needs_grad([src]):
        return ops.segment_sum(src, src, col_a, col_b, width, sign_b, offsets, perm, n_rows)
    return _SegSumFn.apply(src, (col_a, col_b, width, sign_b, offsets, perm, n_rows, i_first, i_second))
