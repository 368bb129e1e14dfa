"""Encoder / GN_Block / decoder data-flows on the CUDA kernels (the product hot path).

One function per reference sub-module; each issues only C-ABI kernel calls (``ops``).  Families and
their block order follow SURVEY.md section 8a:

  fvgn    : vertex segment-sum(e) -> node MLP(x, mean3) -> edge MLP(e, x'[row], x'[col])   Fvgn.py:274-325
  mgn     : edge MLP(e, x[row], x[col]) -> vertex segment-sum(e') -> node MLP               Mgn.py:216-267
  cons_a  : edge MLP(e, x[row]+x[col]) [* asym] -> signed cell segment-sum(e') -> node MLP  Conservative.py:210-254
  cons_e  : edge MLP(e, x[row]+x[col]) -> cell sums of e' (half +/+, half +/-) -> node MLP  Conservative.py:677-732
  cons_f  : vertex sum(sym half) + signed cell sum(asym half) -> node MLP -> edge MLP(e, x'[row], x'[col])  :763-821
  vertpot : fvgn + full-width vertex sum of e'                                              VertPot.py:195-222

The second sub-block consumes the first one's RAW output; both residuals are applied by the
kernels' epilogues (out_sum = residual + out), so no clone / add pass exists.

Inference (no gradient recorded, split tensor-core precision) runs the same data-flows in ``Fast`` mode: the residual
streams x / e are updated IN PLACE (the add is done by the epilogue's TMA reduce-store) and the matrix the edge block
gathers (x' for the fvgn order, x for the mgn order) is handed over as a 16-bit hi|lo split shadow written by the node
block's epilogue, which the edge kernel stages with TMA gather4 - the fp32 copy of x' is never written.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import autograd_ops as A
from . import ops
from ._lib import ACT_SILU, ACT_TANH, PREC_F32, SEG_DIFF2, SEG_DIRECT, SEG_GATHER, SEG_MEAN3, SEG_SUM2, SEG_SUM3S
from .ops import MLPWeights, Seg
from .topology import MeshTopology

H = 128


def _split_mlp(seq: torch.nn.Module):
    """(inner Sequential of Linear/act/Linear/act/Linear, LayerNorm or None)."""
    if isinstance(seq[0], torch.nn.Sequential):
        return seq[0], seq[1]
    return seq, None


def weights_of(seq: torch.nn.Module, act: int = ACT_SILU) -> MLPWeights:
    """MLPWeights view of a reference-layout MLP module, cached on the module and refreshed when a
    parameter is replaced or modified in place (``_version``), e.g. by an optimizer step.

    ``config.training.dropout_rate > 0`` (Model.py:29-33: a Dropout after each SiLU): in training mode the view carries
    ``drop_p`` and the second / third Linear's weights divided by (1 - p) - the kernel drops hidden units by replacing
    their pre-activation (include/gnnfd_b200.h: dropout_p), the rescale of the kept ones is folded into the weights of
    the layer that consumes them and undone on their gradients by ``ops.mlp_backward``.  In eval mode Dropout is the
    identity and the view is the plain one."""
    inner, ln = _split_mlp(seq)
    drops = [m.p for m in inner if isinstance(m, torch.nn.Dropout)]
    drop_p = float(drops[0]) if (seq.training and drops and drops[0] > 0) else 0.0
    if drop_p > 0.0 and (act != ACT_SILU or len(drops) != 2 or drops[1] != drops[0] or drop_p >= 1.0):
        raise NotImplementedError("training-mode dropout is implemented for the reference's layout only: one Dropout of the "
                                  "same rate p < 1 after each of the two SiLUs (Model.py:26-35)")
    lin = [m for m in inner if isinstance(m, torch.nn.Linear)]
    params = [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias]
    if ln is not None:
        params += [ln.weight, ln.bias]
    key = tuple((p.data_ptr(), p._version) if p is not None else None for p in params) + (drop_p,)
    cached = getattr(seq, "_gnnfd_w", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    d = lambda p: None if p is None else p.detach()
    keep = lambda p: d(p) if drop_p == 0.0 else d(p) / (1.0 - drop_p)
    w = MLPWeights(w1=d(lin[0].weight), b1=d(lin[0].bias), w2=keep(lin[1].weight), b2=d(lin[1].bias),
                   w3=keep(lin[2].weight), b3=d(lin[2].bias),
                   ln_w=d(ln.weight) if ln is not None else None,
                   ln_b=d(ln.bias) if ln is not None else None,
                   has_ln=ln is not None, ln_eps=ln.eps if ln is not None else 1e-5, act=act, drop_p=drop_p)
    object.__setattr__(seq, "_gnnfd_w", (key, w))
    return w


# --- sub-blocks -----------------------------------------------------------------------------------

class Fast:
    """Per-processor-run scratch of the inference fast path: the split shadow of the gathered cell matrix."""

    def __init__(self, n_cells: int, prec: int, device):
        self.dtype = ops.split_dtype(prec)
        self.xs = torch.empty(n_cells, 2 * H, dtype=self.dtype, device=device)

    def gather_seg(self, idx) -> Seg:
        """GATHER segment whose source exists ONLY as the split shadow (src aliases it: the C side then insists on the
        TMA path instead of ever reading fp32 rows)."""
        return Seg(self.xs.view(torch.float32), SEG_GATHER, (idx,), split=self.xs)


FAST_ENABLED = True      # tests / comparisons can pin the register-staged kernels with ``no_fast()``
# Conservative 'cons_a' blocks: fold the signed edge->cell sum into the node MLP's input assembly (SEG_SUM3S) instead of a
# separate segment-sum launch.  Correct (tests/test_gpu_parity.py) but MEASURED SLOWER on the 200k-cell rollout (7.88 vs
# 6.76 ms/step): three register-staged 512 B gathers per cell in the node kernel's producers cost more than the segment-sum
# kernel plus one contiguous read of agg[N, 128], so it is off by default.
FUSE_SIGNED_SUM = False


class no_fast:
    """Context manager: run inference through the register-staged (training-shaped) kernels instead of the fast path.
    The two differ in the last bit of the LayerNorm epilogue, so bit-for-bit comparisons against a path that only
    exists in the register-staged form (the peer-memory halo) use this."""

    def __enter__(self):
        global FAST_ENABLED
        self._old, FAST_ENABLED = FAST_ENABLED, False
        return self

    def __exit__(self, *exc):
        global FAST_ENABLED
        FAST_ENABLED = self._old
        return False


def fast_mode(blocks, tensors, prec: int) -> bool:
    """In-place residuals + split shadows are legal when nothing records a gradient and the precision has split
    operands (bf16x3 / fp16x3)."""
    if not FAST_ENABLED or ops.split_dtype(prec) is None:
        return False
    if blocks.training and any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in blocks.modules()):
        return False          # train()-mode forward of a dropout model (even under no_grad): the masking epilogue
    if not torch.is_grad_enabled():
        return True
    return not (any(p.requires_grad for p in blocks.parameters()) or any(t.requires_grad for t in tensors))


def vertex_half_sum(e: torch.Tensor, topo: MeshTopology) -> torch.Tensor:
    """vsum[V, H/2]: half 0 of each face latent onto its first vertex, half 1 onto its second
    (scatter_add of Fvgn.py:312-314) as a deterministic CSR segment sum."""
    return A.segment_sum(e, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices, topo.v0, topo.v1)


def node_mlp_two_hop(seq, x, vsum, topo, prec, want_raw, residual=True, fast: Optional[Fast] = None,
                     split_of_sum: bool = False):
    """x' = cell_mlp(cat[x, (vsum[vf0]+vsum[vf1]+vsum[vf2])/3])  (Fvgn.py:316-323).  ``fast``: x is updated in place and
    the shadow of x' (or of x + x') goes to ``fast.xs``."""
    segs = [Seg(x), Seg(vsum, SEG_MEAN3, topo.vf)]
    if fast is not None:
        return A.mlp(seq, segs, x.shape[0], prec, residual=x, want_raw=want_raw, want_sum=True, inplace=True,
                     out_split=fast.xs, split_of_sum=split_of_sum)
    return A.mlp(seq, segs, x.shape[0], prec, residual=x if residual else None,
                           want_raw=want_raw, want_sum=residual)


def edge_mlp_concat(seq, e, x_src, topo, prec, want_raw, residual=True, fast: Optional[Fast] = None):
    """e' = face_mlp(cat[e, x[row], x[col]])  (Fvgn.py:292-296, Mgn.py:234-238).  ``fast``: e is updated in place and
    x[row] / x[col] are TMA-gathered from the split shadow ``fast.xs`` (``x_src`` is not read)."""
    if fast is not None:
        segs = [Seg(e), fast.gather_seg(topo.row), fast.gather_seg(topo.col)]
        return A.mlp(seq, segs, e.shape[0], prec, residual=e, want_raw=want_raw, want_sum=True, inplace=True)
    segs = [Seg(e), Seg(x_src, SEG_GATHER, (topo.row,)), Seg(x_src, SEG_GATHER, (topo.col,))]
    return A.mlp(seq, segs, e.shape[0], prec, residual=e if residual else None,
                           want_raw=want_raw, want_sum=residual)


def edge_mlp_sum(seq, e, x_src, topo, prec, mul=None, inplace=False):
    """e' = face_mlp(cat[e, x[row] + x[col]]) [* mul]  (Conservative.py:228-234); ``mul`` is ConservativeA's asym
    encoding or ConservativeI's keep matrix (0 on INFLOW / WALL faces, so e + 0 * e' leaves their latent untouched).
    ``inplace`` (inference): e is updated in place - without ``mul`` that is the TMA-store epilogue (raw stored, the residual
    add done by the reduce-store, the residual rows never loaded by the SM)."""
    segs = [Seg(e), Seg(x_src, SEG_SUM2, (topo.row, topo.col))]
    return A.mlp(seq, segs, e.shape[0], prec, mul=mul, residual=e,
                           want_raw=True, want_sum=True, inplace=inplace)


def cell_signed_sum(e_raw: torch.Tensor, topo: MeshTopology) -> torch.Tensor:
    """agg[c] = sum_{col(k)=c} e_k - sum_{row(k)=c} e_k  (Conservative.py:244-249)."""
    off, perm = topo.build_cell_csr()
    return A.segment_sum(e_raw, 0, 0, H, -1.0, off, perm, topo.n_cells, topo.col, topo.row)


def vertex_full_sum(e_raw: torch.Tensor, topo: MeshTopology, n_rows: int) -> torch.Tensor:
    """Vertex_Block (VertPot.py:217-222): full-width sum with ``n_rows`` (= N cells) output rows."""
    return A.segment_sum(e_raw, 0, 0, H, 1.0, topo.vertex_csr_rows(n_rows), topo.vtx_perm, n_rows, topo.v0, topo.v1)


# --- encoder / block / decoder ------------------------------------------------------------------

def mlp_rows(seq, src: torch.Tensor, prec: int, act: int = ACT_SILU) -> torch.Tensor:
    """Plain per-row MLP (encoder / decoder heads)."""
    src = src.contiguous()
    out, _ = A.mlp(seq, [Seg(src)], src.shape[0], prec, act=act)
    return out


def gn_block(family: str, block, x, e, topo: MeshTopology, prec: int = PREC_F32,
             e_asym: Optional[torch.Tensor] = None, want_vertex: bool = False, e_keep: Optional[torch.Tensor] = None,
             fast: Optional[Fast] = None, inplace: bool = False):
    """One GN_Block -> (x_new, e_new, vertex_x or None).  With ``fast`` (inference) x / e are updated in place."""
    if fast is not None and family in ("fvgn", "vertpot"):
        vsum = vertex_half_sum(e, topo)
        cell = block.cell_block.cell_mlp if family == "fvgn" else block.node_block.cell_mlp
        face = block.face_block.face_mlp if family == "fvgn" else block.edge_block.face_mlp
        node_mlp_two_hop(cell, x, vsum, topo, prec, want_raw=False, fast=fast)          # x += x', shadow of x'
        e_raw, _ = edge_mlp_concat(face, e, None, topo, prec, want_raw=want_vertex, fast=fast)      # e += e'
        vx = vertex_full_sum(e_raw, topo, x.shape[0]) if want_vertex else None
        return x, e, vx
    if fast is not None and family == "mgn":
        # fast.xs holds the shadow of the block input x (encoder output, then every node block's x + x')
        e_raw, _ = edge_mlp_concat(block.face_block.face_mlp, e, None, topo, prec, want_raw=True, fast=fast)
        vsum = vertex_half_sum(e_raw, topo)
        node_mlp_two_hop(block.cell_block.cell_mlp, x, vsum, topo, prec, want_raw=False, fast=fast, split_of_sum=True)
        return x, e, None
    if family == "fvgn":
        vsum = vertex_half_sum(e, topo)
        x_raw, x_new = node_mlp_two_hop(block.cell_block.cell_mlp, x, vsum, topo, prec, want_raw=True)
        _, e_new = edge_mlp_concat(block.face_block.face_mlp, e, x_raw, topo, prec, want_raw=False)
        return x_new, e_new, None
    if family == "mgn":
        e_raw, e_new = edge_mlp_concat(block.face_block.face_mlp, e, x, topo, prec, want_raw=True)
        vsum = vertex_half_sum(e_raw, topo)
        _, x_new = node_mlp_two_hop(block.cell_block.cell_mlp, x, vsum, topo, prec, want_raw=False)
        return x_new, e_new, None
    if family == "cons_a" and inplace:
        # inference: both residual streams updated in place through the TMA-store epilogue (see edge_mlp_sum)
        e_raw, _ = edge_mlp_sum(block.face_block.face_mlp, e, x, topo, prec, mul=e_asym, inplace=True)
        agg = cell_signed_sum(e_raw, topo)
        A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(agg)], x.shape[0], prec, residual=x, want_raw=False, want_sum=True,
              inplace=True)
        return x, e, None
    if family == "cons_a":
        e_raw, e_new = edge_mlp_sum(block.face_block.face_mlp, e, x, topo, prec, mul=e_asym)
        ell = getattr(topo, "signed_ell", None)
        if (FUSE_SIGNED_SUM and ell is not None and FAST_ENABLED
                and not (torch.is_grad_enabled() and (e_raw.requires_grad or x.requires_grad))):
            # inference: the signed edge->cell sum is the node MLP's own input assembly (three signed gathered rows per
            # cell, SEG_SUM3S) - agg[N, 128] is never written
            _, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(e_raw, SEG_SUM3S, ell)], x.shape[0], prec, residual=x,
                             want_raw=False, want_sum=True)
            return x_new, e_new, None
        agg = cell_signed_sum(e_raw, topo)
        _, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(agg)], x.shape[0], prec, residual=x,
                         want_raw=False, want_sum=True)
        return x_new, e_new, None
    if family == "cons_e":
        # ConservativeE (Conservative.py:677-732): sum-form face block, then the RAW face output aggregated onto cells:
        # first half with equal signs, second half with opposite signs, agg = cat[sym, asym]
        e_raw, e_new = edge_mlp_sum(block.face_block.face_mlp, e, x, topo, prec, inplace=inplace)
        off, perm = topo.build_cell_csr()
        sym = A.segment_sum(e_raw, 0, 0, H // 2, 1.0, off, perm, topo.n_cells, topo.col, topo.row)
        asym = A.segment_sum(e_raw, H // 2, H // 2, H // 2, -1.0, off, perm, topo.n_cells, topo.col, topo.row)
        _, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(sym), Seg(asym)], x.shape[0], prec, residual=x,
                         want_raw=False, want_sum=True, inplace=inplace)   # cat[x, sym, asym] as three 64-multiple segments
        return x_new, e_new, None
    if family == "cons_f":
        # ConservativeF (Conservative.py:763-821): symmetric half two-hop via the vertices (the same half onto both
        # vertices), antisymmetric half signed edge->cell; then the concat-form face block on the RAW cell output
        vsum = A.segment_sum(e, 0, 0, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices, topo.v0, topo.v1)
        off, perm = topo.build_cell_csr()
        asym = A.segment_sum(e, H // 2, H // 2, H // 2, -1.0, off, perm, topo.n_cells, topo.col, topo.row)
        x_raw, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(vsum, SEG_MEAN3, topo.vf), Seg(asym)],
                             x.shape[0], prec, residual=x, want_raw=True, want_sum=True, inplace=inplace)
        if inplace:
            segs = [Seg(e), Seg(x_raw, SEG_GATHER, (topo.row,)), Seg(x_raw, SEG_GATHER, (topo.col,))]
            _, e_new = A.mlp(block.face_block.face_mlp, segs, e.shape[0], prec, residual=e, want_raw=False, want_sum=True,
                             inplace=True)
        else:
            _, e_new = edge_mlp_concat(block.face_block.face_mlp, e, x_raw, topo, prec, want_raw=False)
        return x_new, e_new, None
    if family in ("cons_g", "cons_i"):
        # ConservativeG / I (Conservative.py:834-896, 1250-1317): F's hybrid cell block, then the SUM-form face block on
        # the raw cell output; I keeps the previous latent on boundary-condition faces (e_keep = 0 rows)
        vsum = A.segment_sum(e, 0, 0, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices, topo.v0, topo.v1)
        off, perm = topo.build_cell_csr()
        asym = A.segment_sum(e, H // 2, H // 2, H // 2, -1.0, off, perm, topo.n_cells, topo.col, topo.row)
        x_raw, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(vsum, SEG_MEAN3, topo.vf), Seg(asym)],
                             x.shape[0], prec, residual=x, want_raw=True, want_sum=True, inplace=inplace)
        _, e_new = edge_mlp_sum(block.face_block.face_mlp, e, x_raw, topo, prec, mul=e_keep if family == "cons_i" else None,
                                inplace=inplace)
        return x_new, e_new, None
    if family == "vertpot":
        vsum = vertex_half_sum(e, topo)
        x_raw, x_new = node_mlp_two_hop(block.node_block.cell_mlp, x, vsum, topo, prec, want_raw=True)
        e_raw, e_new = edge_mlp_concat(block.edge_block.face_mlp, e, x_raw, topo, prec, want_raw=want_vertex)
        vx = vertex_full_sum(e_raw, topo, x.shape[0]) if want_vertex else None
        return x_new, e_new, vx
    raise ValueError(f"unknown family {family!r}")


def gn_block_dual(block, x, e_s, e_a, topo: MeshTopology, prec: int = PREC_F32, inplace: bool = False):
    """ConservativeD GN_Block (Conservative.py:572-645): symmetric and antisymmetric edge streams.
    -> (x_new, e_s_new, e_a_new).  ``inplace`` (inference, ``fast_mode``): the three residual streams are advanced in place
    through the TMA-store epilogue (both face blocks read x before the cell block updates it)."""
    s_raw, s_new = edge_mlp_sum(block.face_block_symm.face_mlp, e_s, x, topo, prec, inplace=inplace)
    segs = [Seg(e_a), Seg(x, SEG_DIFF2, (topo.row, topo.col))]                       # cat[e_a, x[row] - x[col]]
    a_raw, a_new = A.mlp(block.face_block_asym.face_mlp, segs, e_a.shape[0], prec, act=ACT_TANH, residual=e_a,
                         want_raw=True, want_sum=True, inplace=inplace)
    off, perm = topo.build_cell_csr()
    sym = A.segment_sum(s_raw, 0, 0, H, 1.0, off, perm, topo.n_cells, topo.col, topo.row)      # equal signs on both cells
    asym = A.segment_sum(a_raw, 0, 0, H, -1.0, off, perm, topo.n_cells, topo.col, topo.row)    # opposite signs
    _, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(sym), Seg(asym)], x.shape[0], prec, residual=x,
                     want_raw=False, want_sum=True, inplace=inplace)
    return x_new, s_new, a_new


def gn_block_dual_two_hop(block, x, e_s, e_a, topo: MeshTopology, prec: int = PREC_F32, inplace: bool = False):
    """ConservativeH / J GN_Block (Conservative.py:1098-1184), cell block first.  -> (x_new, e_s_new, e_a_new).
    ``inplace``: see ``gn_block_dual`` (the face blocks gather the cell block's RAW output, a separate matrix)."""
    vsum = A.segment_sum(e_s, 0, 0, H, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices, topo.v0, topo.v1)
    off, perm = topo.build_cell_csr()
    asym = A.segment_sum(e_a, 0, 0, H, -1.0, off, perm, topo.n_cells, topo.col, topo.row)
    x_raw, x_new = A.mlp(block.cell_block.cell_mlp, [Seg(x), Seg(vsum, SEG_MEAN3, topo.vf), Seg(asym)], x.shape[0], prec,
                         residual=x, want_raw=True, want_sum=True, inplace=inplace)
    _, s_new = A.mlp(block.face_block_symm.face_mlp, [Seg(e_s), Seg(x_raw, SEG_SUM2, (topo.row, topo.col))],
                     e_s.shape[0], prec, residual=e_s, want_raw=False, want_sum=True, inplace=inplace)
    _, a_new = A.mlp(block.face_block_asym.face_mlp, [Seg(e_a), Seg(x_raw, SEG_DIFF2, (topo.row, topo.col))],
                     e_a.shape[0], prec, act=ACT_TANH, residual=e_a, want_raw=False, want_sum=True, inplace=inplace)
    return x_new, s_new, a_new


def run_processor(family: str, blocks, x, e, topo, prec: int = PREC_F32, e_asym=None, hook=None, e_keep=None,
                  fast: Optional[Fast] = None):
    """All GN_Blocks.  VertPot's vertex sum is only live after the last block (VertPot.py:208: it is
    overwritten every block and never fed back), so it is computed once.

    ``fast`` (see ``Fast`` / ``fast_mode``): x and e - the encoder outputs - are updated in place; for the mgn order
    ``fast.xs`` must already hold the split shadow of x (``encode_cells``)."""
    vx = None
    n = len(blocks)
    # Conservative families in inference: x, e are the encoder's fresh outputs and are advanced in place (no shadow: the
    # sum-form face block gathers fp32 rows)
    inplace = (family in ("cons_a", "cons_e", "cons_f", "cons_g", "cons_i") and fast is None and x.shape[0] > 0
               and not (family == "cons_a" and FUSE_SIGNED_SUM)
               and fast_mode(blocks, [x, e] + [t for t in (e_asym, e_keep) if t is not None], prec))
    for i, blk in enumerate(blocks):
        x, e, vx_i = gn_block(family, blk, x, e, topo, prec,
                              e_asym=e_asym if (family == "cons_a" and i == 0) else None,
                              want_vertex=(family == "vertpot" and i == n - 1), e_keep=e_keep, fast=fast, inplace=inplace)
        if vx_i is not None:
            vx = vx_i
        if hook is not None:
            hook(i, x.clone(), e.clone()) if (fast is not None or inplace) else hook(i, x, e)
    return x, e, vx


FAST_FAMILIES = ("fvgn", "vertpot", "mgn")


def encode_cells(seq, c_x: torch.Tensor, prec: int, family: str, blocks, n_cells: int):
    """Cell encoder + the fast-path decision for the processor that follows -> (x0, Fast or None).  For the mgn order
    the encoder's epilogue also writes the split shadow of x0 that the first edge block gathers."""
    fast = None
    if family in FAST_FAMILIES and fast_mode(blocks, [c_x], prec) and c_x.shape[0] > 0:
        fast = Fast(n_cells, prec, c_x.device)
    src = c_x.contiguous()
    if fast is not None and family == "mgn":
        x0, _ = A.mlp(seq, [Seg(src)], src.shape[0], prec, out_split=fast.xs)
        return x0, fast
    return mlp_rows(seq, src, prec), fast
