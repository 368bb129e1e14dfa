"""Training path of the hot path: encoder -> GN_Blocks -> decoder as ONE autograd node whose backward
is a hand-scheduled chain of CUDA kernels (BASELINE.json north_star (e)).

The reference trains through autograd over ~25 library kernels per block (``src/train.py:253-256``);
here the forward is the same fused kernels as inference (with the backward stash switched on) and
the backward of every sub-block is:

  LayerNorm backward (streaming, + d ln_w / d ln_b / d b3 column sums)      ops.ln_backward
  dgrad chain  dA2 = (dy W3) * act'(a2), dA1 = (dA2 W2) * act'(a1), dIn = dA1 W1[:, segment]
               - tcgen05 single-Linear launches, activation derivative and gradient accumulation
                 (out = residual + ...) fused in the epilogue                    ops.linear_tc
  wgrad        dW = dA^T H on tcgen05 (split-bf16, MN-major operands), H = act(a) / the gathered MLP input assembled on the
               fly, bias gradients as column sums of dA                          ops.wgrad
  transposed gathers: the Face_Block's x[row], x[col] gathers become a deterministic segment sum over
               the CSR of cat[row; col]; the 3-vertex mean becomes a segment sum over cat[vf0; vf1; vf2];
               the edge->vertex scatter_add becomes a gather                     ops.segment_sum3 / gather_pair_add

Families: 'fvgn' (Fvgn/Flux), 'mgn' (Mgn/StreamFunc) and 'vertpot' (VertPot: FVGN blocks + the Vertex_Block on the
last block's raw edge output + a second decoder head).  No atomics anywhere: gradients are bitwise
reproducible run to run.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import ACT_SILU, SEG_GATHER, SEG_MEAN3
from .ops import MLPStash, MLPWeights, Seg
from .processor import H, _split_mlp, vertex_half_sum, weights_of
from .topology import MeshTopology

PARAM_NAMES = ("w1", "b1", "w2", "b2", "w3", "b3", "ln_w", "ln_b")


class Site:
    """One MLP module of the stack and its parameters in PARAM_NAMES order (None where absent)."""

    def __init__(self, seq: torch.nn.Module, act: int = ACT_SILU):
        self.seq, self.act = seq, act
        inner, ln = _split_mlp(seq)
        lin = [m for m in inner if isinstance(m, torch.nn.Linear)]
        self.params = [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias,
                       ln.weight if ln is not None else None, ln.bias if ln is not None else None]

    def weights(self) -> MLPWeights:
        return weights_of(self.seq, self.act)


def mlp_backward(w: MLPWeights, st: MLPStash, segs: Sequence[Seg], rows: int, g: torch.Tensor, prec: int,
                 din: Sequence[Optional[dict]], workspace: torch.Tensor):
    """Backward of one fused MLP through the library's single-call schedule (gnnfd_mlp_backward)."""
    grads, dins = ops.mlp_backward(segs, w, st, rows, g, prec, din, workspace)
    return [grads.get(n) for n in PARAM_NAMES], dins


def edge_mlp_backward_node_side(w: MLPWeights, st: MLPStash, e_in, x_src, topo: MeshTopology, g, d_e_res, d_x_base,
                                prec: int, workspace: torch.Tensor):
    """Backward of the concat-form face MLP  e' = MLP(cat[e, x[row], x[col]])  with the gathered part done on the NODE
    side.  By linearity  sum_k dA1[k]^T x[row[k]] = S_row^T x  and  sum_{row[k]=n} dA1[k] W = S_row[n] W  with
    S_row = segment-sum of dA1 by row (likewise col): dA1 is reduced onto the cells once (two deterministic CSR sums
    onto [N, 128] matrices) and the weight-gradient GEMMs of the two gathered column blocks and the input-gradient
    Linear then run over N contiguous node rows instead of E gathered edge rows, with no [E, 128] temporaries.
    Returns (parameter gradients in PARAM_NAMES order, d_e = d_e_res + dIn_0, d_x = d_x_base + dIn_row + dIn_col)."""
    E, N = e_in.shape[0], x_src.shape[0]
    dev = g.device
    d_a1 = torch.empty(E, H, dtype=torch.float32, device=dev)
    segs = _edge_segs(e_in, x_src, topo)
    # (skip_wgrad_l1 = 2: dW1[:, 0:128] = dA1^T e and db1 are computed by the call, in the same launch as dW2 / dW3; the
    #  columns of the two gathered segments are left to the node-side GEMM below)
    grads, dins = ops.mlp_backward(segs, w, st, E, g, prec, [{"residual": d_e_res} if d_e_res is not None else {}, None, None],
                                   workspace, da1_out=d_a1, skip_wgrad_l1=2)
    w1g = grads["w1"]
    # S = [S_row | S_col]: dA1 reduced onto the cells by ONE deterministic CSR sum (2 N virtual rows, see the topology)
    rc_off, rc_perm = topo.build_row_col_interleaved_csr()
    s_rc = torch.empty(N, 2 * H, dtype=torch.float32, device=dev)
    ops.segment_sum(d_a1, d_a1, 0, 0, H, 1.0, rc_off, rc_perm, 2 * N, out=s_rc.view(2 * N, H))
    # dW1[:, 128:256] = S_row^T x and dW1[:, 256:384] = S_col^T x as ONE weight-gradient GEMM x^T S (x is read once; N
    # contiguous rows), stored transposed into a [256, 128] scratch whose halves are the two column blocks of dW1
    t_rc = torch.empty(2 * H, H, dtype=torch.float32, device=dev)
    ops.wgrad(Seg(x_src), [Seg(s_rc, col=0, width=H), Seg(s_rc, col=H, width=H)], N, t_rc, transpose_out=True,
              workspace=workspace)
    w1g[:, H:2 * H].copy_(t_rc[:H])
    w1g[:, 2 * H:3 * H].copy_(t_rc[H:])
    # d x = base + S_row W1[:, 128:256] + S_col W1[:, 256:384] = base + S . Wcat with Wcat = [W1[:, 128:256]; W1[:, 256:384]]
    # stacked along K: ONE K = 256 Linear over N rows
    if w.bwd_packs is None:
        w.bwd_packs = {}
    packs = w.bwd_packs.setdefault("node_side", {})
    wcat = packs.get("wcat")
    if wcat is None:
        wcat = packs["wcat"] = torch.cat([w.w1[:, H:2 * H], w.w1[:, 2 * H:3 * H]], dim=0).contiguous()     # [256 (k), 128 (n)]
    d_x = ops.linear_tc(Seg(s_rc), N, wcat, 1, H, H, 2 * H, packs, "w1t_rowcol", prec, residual=d_x_base)
    return [grads.get(n) for n in PARAM_NAMES], dins[0], d_x


def mlp_backward_stepwise(w: MLPWeights, st: MLPStash, segs: Sequence[Seg], rows: int, g: torch.Tensor, prec: int,
                          din: Sequence[Optional[dict]], workspace: Optional[torch.Tensor] = None):
    """The same chain issued kernel by kernel from Python (used by the tests to pin the fused call).  ``g`` = gradient w.r.t. the MLP(+LN) output.  ``din[i]`` is None (segment i
    needs no gradient) or a dict with optional ``residual`` (added to the segment's input gradient) and
    ``out`` (destination).  Returns (parameter gradients in PARAM_NAMES order, [dIn_i or None])."""
    packs = {}
    code = w.act + 1                      # SiLU -> 1, tanh -> 2 (wgrad act / dgrad mul_mode)
    dev = g.device
    n_out, k_in = w.w3.shape[0], w.w1.shape[1]
    grads = {}
    g = g.contiguous()
    if w.has_ln:
        dy, sums = ops.ln_backward(g, st.xhat, st.rstd, w.ln_w)
        if w.ln_w is not None:
            grads["ln_w"], grads["ln_b"] = sums[0], sums[1]
        db3_from_ln = sums[2]
    else:
        dy, db3_from_ln = g, None
    new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    # ---- layer 3
    grads["w3"] = new(n_out, H)
    want_b3 = w.b3 is not None and db3_from_ln is None
    if want_b3:
        grads["b3"] = new(n_out)
    elif w.b3 is not None:
        grads["b3"] = db3_from_ln
    if n_out == H:
        ops.wgrad(Seg(dy), [Seg(st.a2)], rows, grads["w3"], b_act=code,
                  colsum=grads["b3"] if want_b3 else None, workspace=workspace)
    else:   # narrow head: put the 128-wide activation on the M side, store transposed
        ops.wgrad(Seg(st.a2), [Seg(dy)], rows, grads["w3"], a_act=code, transpose_out=True,
                  colsum=grads["b3"] if want_b3 else None, colsum_of_b=True, workspace=workspace)
    d_a2 = ops.linear_tc(Seg(dy), rows, w.w3, 1, H, H, n_out, packs, "w3t", prec, mul=st.a2, mul_mode=code)
    # ---- layer 2
    grads["w2"] = new(H, H)
    if w.b2 is not None:
        grads["b2"] = new(H)
    ops.wgrad(Seg(d_a2), [Seg(st.a1)], rows, grads["w2"], b_act=code, colsum=grads.get("b2"), workspace=workspace)
    d_a1 = ops.linear_tc(Seg(d_a2), rows, w.w2, 1, H, H, H, packs, "w2t", prec, mul=st.a1, mul_mode=code)
    # ---- layer 1
    grads["w1"] = new(H, k_in)
    if w.b1 is not None:
        grads["b1"] = new(H)
    ops.wgrad(Seg(d_a1), list(segs), rows, grads["w1"], colsum=grads.get("b1"), workspace=workspace)
    dins: List[Optional[torch.Tensor]] = []
    col0 = 0
    for i, s in enumerate(segs):
        width = s.width if s.width is not None else s.src.shape[1] - s.col
        spec = din[i] if i < len(din) else None
        if spec is None:
            dins.append(None)
        else:
            dins.append(ops.linear_tc(Seg(d_a1), rows, w.w1[:, col0:], 1, k_in, width, H, packs, ("w1t", col0), prec,
                                      residual=spec.get("residual"), out=spec.get("out")))
        col0 += width
    if w.drop_p > 0.0:      # see ops.mlp_backward
        grads["w2"].mul_(1.0 / (1.0 - w.drop_p))
        grads["w3"].mul_(1.0 / (1.0 - w.drop_p))
    return [grads.get(n) for n in PARAM_NAMES], dins


class Plan:
    """The MLP sites of one model in execution order + the data-flow family."""

    def __init__(self, family: str, enc_edge: Site, enc_node: Site, blocks: Sequence[tuple], dec: Site,
                 dec_vertex: Optional[Site] = None):
        if family not in ("fvgn", "mgn", "vertpot"):
            raise NotImplementedError(f"training kernels cover the 'fvgn', 'mgn' and 'vertpot' families, not {family!r}")
        self.family, self.enc_edge, self.enc_node, self.blocks, self.dec = family, enc_edge, enc_node, list(blocks), dec
        self.dec_vertex = dec_vertex      # VertPot: second head on the vertex sums of the last block's raw edge output
        self.sites = [enc_edge, enc_node] + [s for b in self.blocks for s in b] + [dec]   # block = (node site, edge site)
        if dec_vertex is not None:
            self.sites.append(dec_vertex)

    def flat_params(self):
        return [p for s in self.sites for p in s.params if p is not None]


def _node_segs(x, vsum, topo):
    return [Seg(x), Seg(vsum, SEG_MEAN3, topo.vf)]


def _edge_segs(e, xs, topo):
    return [Seg(e), Seg(xs, SEG_GATHER, (topo.row,)), Seg(xs, SEG_GATHER, (topo.col,))]


class EncodeProcessDecode(torch.autograd.Function):
    """decoder(processor(encoder(c_x, f_x))) with a kernel-scheduled backward; differentiable w.r.t. every
    MLP parameter (inputs are data: no gradient)."""

    @staticmethod
    def forward(ctx, plan: Plan, topo: MeshTopology, prec: int, c_x: torch.Tensor, f_x: torch.Tensor, *params):
        fam = "fvgn" if plan.family == "vertpot" else plan.family     # VertPot blocks are FVGN blocks (VertPot.py:195-210)
        vertpot = plan.family == "vertpot"
        c_x, f_x = c_x.contiguous(), f_x.contiguous()
        N, E = c_x.shape[0], f_x.shape[0]
        e, _, st_ee = ops.mlp_forward([Seg(f_x)], plan.enc_edge.weights(), E, prec, stash=True)
        x, _, st_en = ops.mlp_forward([Seg(c_x)], plan.enc_node.weights(), N, prec, stash=True)
        saved = []
        e_raw_last = None
        for bi, (node_site, edge_site) in enumerate(plan.blocks):
            wn, we = node_site.weights(), edge_site.weights()
            if fam == "fvgn":
                vsum = vertex_half_sum(e, topo)
                x_raw, x_new, st_n = ops.mlp_forward(_node_segs(x, vsum, topo), wn, N, prec, residual=x,
                                                     want_raw=True, want_sum=True, stash=True)
                last_vp = vertpot and bi == len(plan.blocks) - 1
                e_raw_last, e_new, st_e = ops.mlp_forward(_edge_segs(e, x_raw, topo), we, E, prec, residual=e,
                                                          want_raw=last_vp, want_sum=True, stash=True)
                saved.append((x, e, vsum, x_raw, st_n, st_e))
            else:
                e_raw, e_new, st_e = ops.mlp_forward(_edge_segs(e, x, topo), we, E, prec, residual=e,
                                                     want_raw=True, want_sum=True, stash=True)
                vsum = vertex_half_sum(e_raw, topo)
                _, x_new, st_n = ops.mlp_forward(_node_segs(x, vsum, topo), wn, N, prec, residual=x,
                                                 want_raw=False, want_sum=True, stash=True)
                saved.append((x, e, vsum, None, st_n, st_e))
            x, e = x_new, e_new
        dec_in = e if fam == "fvgn" else x
        out, _, st_d = ops.mlp_forward([Seg(dec_in)], plan.dec.weights(), dec_in.shape[0], prec, stash=True)
        ctx.plan, ctx.topo, ctx.prec = plan, topo, prec
        ctx.stash = (c_x, f_x, st_ee, st_en, saved, dec_in, st_d)
        if vertpot:
            # Vertex_Block (VertPot.py:217-222): full-width sum of the last block's RAW edge output, N output rows
            vx = ops.segment_sum(e_raw_last, e_raw_last, 0, 0, H, 1.0, topo.vertex_csr_rows(N), topo.vtx_perm, N)
            out_v, _, st_dv = ops.mlp_forward([Seg(vx)], plan.dec_vertex.weights(), N, prec, stash=True)
            ctx.vertpot = (vx, st_dv)
            return out, out_v
        return out

    @staticmethod
    def backward(ctx, g_out, g_out_v=None):
        plan, topo, prec = ctx.plan, ctx.topo, ctx.prec
        c_x, f_x, st_ee, st_en, saved, dec_in, st_d = ctx.stash
        fam = "fvgn" if plan.family == "vertpot" else plan.family
        N, E, V = c_x.shape[0], f_x.shape[0], topo.n_vertices
        dev = g_out.device
        ws = ops.mlp_backward_workspace(max(N, E), dev)
        vf_off, vf_perm = topo.build_vf_csr()
        site_grads = {}

        gd, dins = mlp_backward(plan.dec.weights(), st_d, [Seg(dec_in)], dec_in.shape[0], g_out, prec, [{}], ws)
        site_grads[id(plan.dec)] = gd
        d_e, d_x = (dins[0], None) if fam == "fvgn" else (None, dins[0])
        g_extra = None          # VertPot: gradient reaching the last block's raw edge output through the vertex head
        if plan.family == "vertpot":
            vx, st_dv = ctx.vertpot
            gv, dins = mlp_backward(plan.dec_vertex.weights(), st_dv, [Seg(vx)], N, g_out_v, prec, [{}], ws)
            site_grads[id(plan.dec_vertex)] = gv
            # transpose of the full-width edge->vertex sum: d e_raw[k] = d vx[v0[k]] + d vx[v1[k]], added to d e_new
            g_extra = ops.gather_pair_add(dins[0], topo.v0, topo.v1, 1.0, False, E, base=d_e)

        for (node_site, edge_site), (x_in, e_in, vsum, x_raw, st_n, st_e) in zip(reversed(plan.blocks), reversed(saved)):
            wn, we = node_site.weights(), edge_site.weights()
            if fam == "fvgn":
                # e_new = e + edge(e, x_raw[row], x_raw[col]);  x_new = x + x_raw;  x_raw = node(x, mean3(S(e)))
                g_edge, g_extra = (g_extra, None) if g_extra is not None else (d_e, None)
                ge, d_e_acc, d_x_raw = edge_mlp_backward_node_side(we, st_e, e_in, x_raw, topo, g_edge, d_e, d_x, prec, ws)
                gn, dins = mlp_backward(wn, st_n, _node_segs(x_in, vsum, topo), N, d_x_raw, prec,
                                        [{"residual": d_x} if d_x is not None else {}, {}], ws)
                d_x, t3 = dins
                d_vsum = ops.segment_sum3(t3, None, None, (0, 0, 0), H // 2, 1.0, N, vf_off, vf_perm, V, scale=1.0 / 3.0)
                d_e = ops.gather_pair_add(d_vsum, topo.v0, topo.v1, 1.0, True, E, base=d_e_acc, out=d_e_acc)
            else:
                # e_raw = edge(e, x[row], x[col]); e_new = e + e_raw;  x_new = x + node(x, mean3(S(e_raw)))
                gn, dins = mlp_backward(wn, st_n, _node_segs(x_in, vsum, topo), N, d_x, prec,
                                        [{"residual": d_x}, {}], ws)
                d_x_acc, t3 = dins
                d_vsum = ops.segment_sum3(t3, None, None, (0, 0, 0), H // 2, 1.0, N, vf_off, vf_perm, V, scale=1.0 / 3.0)
                d_e_raw = ops.gather_pair_add(d_vsum, topo.v0, topo.v1, 1.0, True, E, base=d_e)
                ge, d_e, d_x = edge_mlp_backward_node_side(we, st_e, e_in, x_in, topo, d_e_raw, d_e, d_x_acc, prec, ws)
            site_grads[id(node_site)], site_grads[id(edge_site)] = gn, ge

        if d_e is None:
            d_e = torch.zeros(E, H, dtype=torch.float32, device=dev)
        if d_x is None:
            d_x = torch.zeros(N, H, dtype=torch.float32, device=dev)
        site_grads[id(plan.enc_edge)], _ = mlp_backward(plan.enc_edge.weights(), st_ee, [Seg(f_x)], E, d_e, prec, [None], ws)
        site_grads[id(plan.enc_node)], _ = mlp_backward(plan.enc_node.weights(), st_en, [Seg(c_x)], N, d_x, prec, [None], ws)
        ctx.stash = ctx.vertpot = None
        flat = []
        for s in plan.sites:
            gs = site_grads[id(s)]
            flat.extend(gr for gr, p in zip(gs, s.params) if p is not None)
        return (None, None, None, None, None, *flat)


def encode_process_decode_train(plan: Plan, topo: MeshTopology, prec: int, c_x, f_x):
    return EncodeProcessDecode.apply(plan, topo, prec, c_x, f_x, *plan.flat_params())
