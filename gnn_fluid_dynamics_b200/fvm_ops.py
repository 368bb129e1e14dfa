"""Finite-volume glue as fused kernels with hand-written backward passes.

The reference computes the integrator (``src/models/Fvgn.py:221-255``), the face-area normalisation
(``src/utils/normalisation.py:325-344``), the divergence (``src/utils/fvm.py:26-37``) and the masked MSE terms
(``src/utils/loss.py:55-60``) with ~20 small tensor kernels per forward, boolean-mask indexing (host syncs) and
sort-based indexing backward passes.  Here each is one ``torch.autograd.Function`` = one kernel forward, one kernel
backward, fixed-degree gathers (3 faces per cell, <= 2 cells per face), deterministic reductions; nothing
synchronises with the host, so a whole training step or rollout step stays asynchronous / graph-capturable.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import ops
from ._lib import check, lib

_WS = {}


def _workspace(device) -> torch.Tensor:
    """Reduction scratch of the glue kernels (ticket + block partials), one per device, zero-initialised once: every
    kernel leaves the ticket at zero."""
    ws = _WS.get(device)
    if ws is None:
        ws = torch.zeros(lib.gnnfd_glue_workspace_bytes(), dtype=torch.uint8, device=device)
        _WS[device] = ws
    return ws


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError(f"{name}: expected a CUDA fp32 tensor (this path has no CPU fallback)")
    if t.device.index != ops._current_device():
        raise RuntimeError(f"{name}: tensor is on cuda:{t.device.index}, current device is cuda:{torch.cuda.current_device()}")
    return t


def _rows2d(t: torch.Tensor, name: str):
    """(tensor, row stride) of a [R, C] matrix whose columns are contiguous (column slices of a wider matrix are fine)."""
    _f32(t, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous CUDA int32 tensor")
    return t


def cell_faces(topo, f_face: torch.Tensor):
    """``f_graph.face`` [3, N] (the 3 face ids of each cell) narrowed to int32 once per topology."""
    cf = getattr(topo, "cf", None)
    if cf is None or cf[0].shape[0] != f_face.shape[1]:
        n = ops.index_narrow(f_face, topo.n_faces)
        topo._flags.append(n._gnnfd_range_flag)
        cf = (n[0], n[1], n[2])
        topo.cf = cf
    return cf


# ------------------------------------------------------------------------------------- face-area BatchNorm
class _FaceAreaNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, area, volume, row, col, dt, bn, n_updates):
        E = area.shape[0]
        dev = area.device
        out = torch.empty(E, 1, dtype=torch.float32, device=dev)
        training = bool(bn.training or not bn.track_running_stats)
        stats = torch.empty(2, dtype=torch.float32, device=dev) if training else None
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        ws = _workspace(dev)
        nbt = bn.num_batches_tracked
        check(lib.gnnfd_face_area_norm(
            area.data_ptr(), volume.data_ptr(), row.data_ptr(), col.data_ptr(), dt.data_ptr(), dt.numel(), E,
            None if weight is None else weight.data_ptr(), None if bias is None else bias.data_ptr(),
            bn.running_mean.data_ptr(), bn.running_var.data_ptr(), None if nbt is None else nbt.data_ptr(),
            int(training), momentum, float(bn.eps), int(n_updates), out.data_ptr(),
            None if stats is None else stats.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream()), "gnnfd_face_area_norm")
        ops._count(2 if training else 1)
        ctx.save_for_backward(area, volume, row, col, dt, bn.running_mean, bn.running_var)
        ctx.stats, ctx.eps, ctx.has = stats, float(bn.eps), (weight is not None, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        area, volume, row, col, dt, rm, rv = ctx.saved_tensors
        dev = area.device
        dwb = torch.empty(2, dtype=torch.float32, device=dev)
        g = g.contiguous()
        ws = _workspace(dev)
        check(lib.gnnfd_face_area_norm_backward(
            area.data_ptr(), volume.data_ptr(), row.data_ptr(), col.data_ptr(), dt.data_ptr(), dt.numel(), area.shape[0],
            None if ctx.stats is None else ctx.stats.data_ptr(), rm.data_ptr(), rv.data_ptr(), ctx.eps, g.data_ptr(),
            dwb.data_ptr(), dwb.data_ptr() + 4, ws.data_ptr(), ws.numel(), ops._stream()), "gnnfd_face_area_norm_backward")
        ops._count(1)
        return (dwb[0:1] if ctx.has[0] else None, dwb[1:2] if ctx.has[1] else None,
                None, None, None, None, None, None, None)


def face_area_norm(f_area: torch.Tensor, c_volume: torch.Tensor, row: torch.Tensor, col: torch.Tensor, dt: torch.Tensor,
                   bn: torch.nn.BatchNorm1d, n_updates: int = 1) -> torch.Tensor:
    """``BatchNorm1d(1)(face_area * mean(dt) / ((vol[row] + vol[col]) / 2))`` -> [E, 1]
    (normalisation.py:325-344).  Batch statistics and the running-stat update (applied ``n_updates`` times: the
    reference normalises once in the integrator and once in the loss of the same step) when ``bn.training``."""
    area = _f32(f_area, "face_area").reshape(-1).contiguous()
    vol = _f32(c_volume, "cell_volume").reshape(-1).contiguous()
    dt = _f32(dt, "dt").reshape(-1).contiguous()
    return _FaceAreaNorm.apply(bn.weight, bn.bias, area, vol, _i32(row, "row"), _i32(col, "col"), dt, bn, n_updates)


# --------------------------------------------------------------------------------------- integrator / divergence
class _FvmIntegrate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, edge_out, area, normal, cf0, cf1, cf2, row, col, rho, want_acc, want_div):
        eo, ld = _rows2d(edge_out, "edge_out")
        N = normal.shape[0]
        dev = eo.device
        acc = torch.empty(N, 2, dtype=torch.float32, device=dev) if want_acc else None
        div = torch.empty(N, 1, dtype=torch.float32, device=dev) if want_div else None
        check(lib.gnnfd_fvm_integrate(eo.data_ptr(), ld, area.data_ptr(), normal.data_ptr(), cf0.data_ptr(), cf1.data_ptr(),
                                      cf2.data_ptr(), N, float(rho), None if acc is None else acc.data_ptr(),
                                      None if div is None else div.data_ptr(), ops._stream()), "gnnfd_fvm_integrate")
        ops._count(1)
        ctx.save_for_backward(eo, area, normal, cf0, cf1, cf2, row, col)
        ctx.ld, ctx.rho, ctx.shape, ctx.modes = ld, float(rho), tuple(edge_out.shape), (want_acc, want_div)
        if want_acc and want_div:
            return acc, div
        return acc if want_acc else div

    @staticmethod
    def backward(ctx, *grads):
        eo, area, normal, cf0, cf1, cf2, row, col = ctx.saved_tensors
        want_acc, want_div = ctx.modes
        g_acc = grads[0] if want_acc else None
        g_div = grads[1] if (want_acc and want_div) else (grads[0] if want_div else None)
        g_acc = None if g_acc is None else g_acc.contiguous()
        g_div = None if g_div is None else g_div.contiguous()
        E, ncols = ctx.shape[0], ctx.shape[1]
        dev = eo.device
        n_out = 5 if want_acc else 2
        d_eo = torch.empty(E, ncols, dtype=torch.float32, device=dev) if ncols == n_out else \
            torch.zeros(E, ncols, dtype=torch.float32, device=dev)
        d_area = torch.empty(E, dtype=torch.float32, device=dev)      # `area` enters flattened
        check(lib.gnnfd_fvm_integrate_backward(
            eo.data_ptr(), ctx.ld, area.data_ptr(), normal.data_ptr(), cf0.data_ptr(), cf1.data_ptr(), cf2.data_ptr(),
            row.data_ptr(), col.data_ptr(), E, ctx.rho, None if g_acc is None else g_acc.data_ptr(),
            None if g_div is None else g_div.data_ptr(), d_eo.data_ptr(), ncols, n_out, d_area.data_ptr(), ops._stream()),
            "gnnfd_fvm_integrate_backward")
        ops._count(1)
        return d_eo, d_area, None, None, None, None, None, None, None, None, None


def _prep(area, normal, cf, row, col):
    area = _f32(area, "face_area").reshape(-1).contiguous()
    normal = _f32(normal, "cell_normal").contiguous()
    if normal.dim() != 3 or normal.shape[1:] != (3, 2):
        raise RuntimeError(f"cell_normal: expected [N, 3, 2], got {tuple(normal.shape)}")
    return area, normal, [_i32(t, "cell_face") for t in cf], _i32(row, "row"), _i32(col, "col")


def fvm_integrate(edge_out: torch.Tensor, area: torch.Tensor, normal: torch.Tensor, cf: Sequence[torch.Tensor],
                  row: torch.Tensor, col: torch.Tensor, rho: float = 1.0) -> torch.Tensor:
    """acc[c] = -(sum_j u_f (u_f . n_cj) a_f) - (sum_j p_f n_cj a_f) / rho + sum_j d_f over the three faces f = cf[j][c]
    of each cell; ``edge_out`` rows are (u, v, p, d0, d1).  FvgnA Integrator.forward, Fvgn.py:221-255."""
    area, normal, cf, row, col = _prep(area, normal, cf, row, col)
    return _FvmIntegrate.apply(edge_out, area, normal, cf[0], cf[1], cf[2], row, col, rho, True, False)


class _Gather3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, cf0, cf1, cf2, row, col):
        t2, ld = _rows2d(t, "t")
        N, w = cf0.shape[0], t2.shape[1]
        out = torch.empty(3, N, w, dtype=torch.float32, device=t2.device)
        check(lib.gnnfd_gather3(t2.data_ptr(), ld, w, cf0.data_ptr(), cf1.data_ptr(), cf2.data_ptr(), N, out.data_ptr(),
                                ops._stream()), "gnnfd_gather3")
        ops._count(1)
        ctx.save_for_backward(cf0, cf1, cf2, row, col)
        ctx.shape = tuple(t.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        cf0, cf1, cf2, row, col = ctx.saved_tensors
        g = g.contiguous()
        N, w = g.shape[1], g.shape[2]
        E = ctx.shape[0]
        d = torch.empty(E, w, dtype=torch.float32, device=g.device)
        check(lib.gnnfd_gather3_backward(g.data_ptr(), w, cf0.data_ptr(), cf1.data_ptr(), cf2.data_ptr(), row.data_ptr(),
                                         col.data_ptr(), N, E, d.data_ptr(), w, ops._stream()), "gnnfd_gather3_backward")
        ops._count(1)
        return d.view(ctx.shape), None, None, None, None, None


def gather3(t: torch.Tensor, cf: Sequence[torch.Tensor], row: torch.Tensor, col: torch.Tensor) -> torch.Tensor:
    """``stack([t[cf[0]], t[cf[1]], t[cf[2]]])`` -> [3, N, w] for a per-face matrix ``t`` [E, w] (w <= 8): the three
    ``x[f_graph.face[j]]`` gathers of the integrators / divergences as ONE kernel whose autograd is a fixed-degree,
    sort-free transpose (each face collects from its <= 2 cells ``row[f]``, ``col[f]``); see gnnfd_gather3."""
    if t.dim() != 2 or t.shape[1] > 8:
        raise RuntimeError(f"gather3: expected [E, w <= 8], got {tuple(t.shape)}")
    cf = [_i32(x, "cell_face") for x in cf]
    return _Gather3.apply(_f32(t, "t"), cf[0], cf[1], cf[2], _i32(row, "row"), _i32(col, "col"))


def flux_integrate(edge_out: torch.Tensor, coeff: Optional[torch.Tensor], area: Optional[torch.Tensor],
                   normal: Optional[torch.Tensor], cf: Sequence[torch.Tensor], row: torch.Tensor, col: torch.Tensor,
                   rho: float = 1.0, want_acc: bool = True, want_cell_flux: bool = False, flux_col: int = 3,
                   flux_scale: float = 1.0, flux_shift: float = 0.0):
    """FluxA's integrator on the signed per-cell face flux (Flux.py:166-206, fvm.py:96-156) as ONE kernel, forward only:
    ``edge_out`` rows are (u, v, p, phi, d0, d1); returns acc [N, 2] and / or cell_flux [N, 3] =
    (edge_out[:, flux_col] * flux_scale + flux_shift)[cf] * sign (see gnnfd_flux_integrate).  Bit-identical to the tensor
    expression; no autograd (the training path keeps the tensor code)."""
    if torch.is_grad_enabled() and edge_out.requires_grad:
        raise RuntimeError("flux_integrate is forward-only (evaluation / rollout)")
    eo, ld = _rows2d(edge_out, "edge_out")
    cf = [_i32(t, "cell_face") for t in cf]
    row, col = _i32(row, "row"), _i32(col, "col")
    N, dev = cf[0].shape[0], eo.device
    ptr = lambda t: None if t is None else t.data_ptr()
    if want_acc:
        coeff = _f32(coeff, "coeff").reshape(-1).contiguous()
        area = _f32(area, "face_area").reshape(-1).contiguous()
        normal = _f32(normal, "cell_normal").contiguous()
        if normal.dim() != 3 or normal.shape[1:] != (3, 2):
            raise RuntimeError(f"cell_normal: expected [N, 3, 2], got {tuple(normal.shape)}")
    acc = torch.empty(N, 2, dtype=torch.float32, device=dev) if want_acc else None
    cfl = torch.empty(N, 3, dtype=torch.float32, device=dev) if want_cell_flux else None
    check(lib.gnnfd_flux_integrate(eo.data_ptr(), ld, flux_col, ptr(coeff) if want_acc else None,
                                   ptr(area) if want_acc else None, ptr(normal) if want_acc else None, cf[0].data_ptr(),
                                   cf[1].data_ptr(), cf[2].data_ptr(), row.data_ptr(), col.data_ptr(), N, float(rho),
                                   ptr(acc), ptr(cfl), float(flux_scale), float(flux_shift), ops._stream()),
          "gnnfd_flux_integrate")
    ops._count(1)
    if want_acc and want_cell_flux:
        return acc, cfl
    return acc if want_acc else cfl


def fvm_divergence(face_velocity: torch.Tensor, area: torch.Tensor, normal: torch.Tensor, cf: Sequence[torch.Tensor],
                   row: torch.Tensor, col: torch.Tensor) -> torch.Tensor:
    """div[c] = sum_j (u_f . n_cj) a_f  -> [N, 1]   (fvm.py:26-37)."""
    area, normal, cf, row, col = _prep(area, normal, cf, row, col)
    return _FvmIntegrate.apply(face_velocity, area, normal, cf[0], cf[1], cf[2], row, col, 1.0, False, True)


# ------------------------------------------------------------------------------------------------ masked MSE
class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mask):
        a2, ld_a = _rows2d(a, "output")
        b2, ld_b = _rows2d(b, "target")
        R, Cc = a2.shape
        dev = a2.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        ws = _workspace(dev)
        check(lib.gnnfd_masked_mse(a2.data_ptr(), ld_a, b2.data_ptr(), ld_b, None if mask is None else mask.data_ptr(), R, Cc,
                                   out.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream()), "gnnfd_masked_mse")
        ops._count(1)
        ctx.save_for_backward(a2, b2, out, *([mask] if mask is not None else []))
        ctx.lds, ctx.shape = (ld_a, ld_b), tuple(a.shape)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a2, b2, out = ctx.saved_tensors[:3]
        mask = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
        R, Cc = a2.shape
        d = torch.empty(R, Cc, dtype=torch.float32, device=a2.device)
        g = g.reshape(1).contiguous().float()
        check(lib.gnnfd_masked_mse_backward(a2.data_ptr(), ctx.lds[0], b2.data_ptr(), ctx.lds[1],
                                            None if mask is None else mask.data_ptr(), R, Cc, out.data_ptr(), g.data_ptr(),
                                            d.data_ptr(), Cc, ops._stream()), "gnnfd_masked_mse_backward")
        ops._count(1)
        return d.reshape(ctx.shape), None, None


def masked_mse(output: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean((output - target)[mask] ** 2) as one kernel - ``MSE_per_element_torch`` (utils/loss.py:55-60) without the
    boolean-mask indexing (no host sync, no sort-based index backward).  ``mask``: bool [rows] or None."""
    if output.dim() == 1:
        output, target = output.unsqueeze(-1), target.unsqueeze(-1)
    if output.shape != target.shape:
        raise RuntimeError(f"masked_mse: shapes differ {tuple(output.shape)} vs {tuple(target.shape)}")
    if output.dim() != 2:
        output, target = output.reshape(output.shape[0], -1), target.reshape(target.shape[0], -1)
    m = None
    if mask is not None:
        m = mask.reshape(-1)
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        elif m.dtype != torch.uint8:
            m = (m != 0).view(torch.uint8)
        m = m.contiguous()
        if m.shape[0] != output.shape[0]:
            raise RuntimeError("masked_mse: mask length != rows")
    return _MaskedMSE.apply(output, target.detach(), m)


MSE_LOSS_NAMES = ("MSE_per_element_torch", "MSE_per_element", "mse")


def is_plain_mse(loss_func) -> bool:
    """True for the reference's element-wise masked mean-squared-error callables (``train.py:372`` passes
    ``MSE_per_element_torch``), whose arithmetic ``masked_mse`` restates; any other callable is called as given."""
    return getattr(loss_func, "__name__", "") in MSE_LOSS_NAMES


# --------------------------------------------------------------------------------------------- state advance
def state_advance(x_raw: torch.Tensor, delta: torch.Tensor, has_change: bool, row: torch.Tensor, col: torch.Tensor,
                  f_raw: Optional[torch.Tensor], bc_mask: Optional[torch.Tensor], bc_value: Optional[torch.Tensor],
                  x_norm: Optional[torch.Tensor] = None, cell_stats: Optional[Sequence[float]] = None,
                  f_norm: Optional[torch.Tensor] = None, face_stats: Optional[Sequence[float]] = None,
                  vel_out: Optional[torch.Tensor] = None) -> None:
    """One rollout state advance (rollout.py:340 + update_features, Fvgn.py:133-148 / Mgn.py:139-151, + the next step's
    input normalisation, normalisation.py:255-278) in two kernels:
    ``x_raw[:, :2] = vel = x_raw[:, :2] + delta`` (or ``= delta``), ``f_raw[:, :2] = where(bc, bc_value, vel[row] - vel[col])``
    and optionally their z-scored copies ``x_norm`` / ``f_norm`` with (mean0, scale0, mean1, scale1)."""
    x_raw, ld_x = _rows2d(x_raw, "x_raw")
    delta, ld_d = _rows2d(delta, "delta")
    arr4 = lambda s: None if s is None else (C.c_float * 4)(*[float(v) for v in s])
    cs, fs = arr4(cell_stats), arr4(face_stats)
    m = None
    if bc_mask is not None:
        m = bc_mask.reshape(-1)
        m = (m.view(torch.uint8) if m.dtype == torch.bool else m).contiguous()
    ptr = lambda t: None if t is None else t.data_ptr()
    ld = lambda t: 0 if t is None else t.stride(0)
    check(lib.gnnfd_state_advance(
        x_raw.data_ptr(), ld_x, delta.data_ptr(), ld_d, int(has_change), x_raw.shape[0], ptr(x_norm), ld(x_norm), cs,
        ptr(row), ptr(col), ptr(m), ptr(bc_value), ld(bc_value), 0 if f_raw is None else f_raw.shape[0], ptr(f_raw),
        ld(f_raw), ptr(f_norm), ld(f_norm), fs, ptr(vel_out), ops._stream()), "gnnfd_state_advance")
    ops._count(2)
