"""Oracle: encoder / GN_Block / decoder data-flows of the reference's model families, CPU fp32.

Every function takes a reference-layout ``state_dict`` (same keys the reference's modules produce)
plus plain tensors, so the same parameters drive the reference (golden fixtures), this oracle and
the CUDA path.

Families:
  'fvgn'   FvgnA / FluxA-D / FvgnB..K  : Cell_Block -> Face_Block      (Fvgn.py:268-325)
  'mgn'    MgnA-C / StreamFuncA-D      : Face_Block -> Cell_Block      (Mgn.py:210-267)
  'cons_a' ConservativeA / B           : sum-form face block, signed edge->cell sum
                                                                       (Conservative.py:204-254)
  'vertpot' VertPotA..                 : fvgn order + Vertex_Block     (VertPot.py:187-222)
"""
from __future__ import annotations

import torch

from .mlp import mlp_from_state
from .scatter import scatter_add

FAMILIES = ("fvgn", "mgn", "cons_a", "vertpot", "cons_e", "cons_f", "cons_d", "cons_g", "cons_i", "cons_h", "fvgn_f")

_FAMILY_OF = {
    "FvgnA": "fvgn", "FluxA": "fvgn", "MgnA": "mgn", "StreamFuncA": "mgn",
    "ConservativeA": "cons_a", "ConservativeB": "cons_a", "VertPotA": "vertpot",
    "ConservativeE": "cons_e", "ConservativeF": "cons_f", "ConservativeD": "cons_d",
    "ConservativeG": "cons_g", "ConservativeI": "cons_i", "ConservativeH": "cons_h", "ConservativeJ": "cons_h",
    "FvgnF": "fvgn_f", "ConservativeK": "cons_h",      # K = H with a half-width antisymmetric stream (same data-flow)
    # glue-only variants: MgnA's encoder / processor / decoder (Mgn.py:278-424, StreamFunc.py:109-235)
    "VertPotB": "vertpot", "VertPotC": "vertpot", "VertPotE": "vertpot", "VertPotG": "vertpot",
    "ConservativeB": "cons_b", "ConservativeJ": "cons_h",      # J = H's network, other glue (Conservative.py:1320-1683)
    "FvgnB": "fvgn", "FvgnC": "fvgn", "FvgnD": "fvgn", "FvgnE": "fvgn", "FvgnH": "fvgn", "FvgnI": "fvgn", "FvgnJ": "fvgn", "FvgnK": "fvgn",
    "FluxB": "fvgn", "FluxC": "fvgn", "FluxD": "fvgn",      # FvgnA's network, other integrators (Flux.py:209-595)
    "MgnB": "mgn", "MgnC": "mgn", "StreamFuncA": "mgn", "StreamFuncB": "mgn", "StreamFuncC": "mgn", "StreamFuncD": "mgn",
}


def family_of(model_name: str) -> str:
    return _FAMILY_OF[model_name]


# --- sub-blocks -----------------------------------------------------------------------------

def two_hop_aggregate(e, v_edge_index, v_face, n_vertices):
    """Cell_Block aggregation (Fvgn.py:305-321 == Mgn.py:247-263).

    The edge latent is split in halves; half 0 is summed onto the face's first vertex
    (v_edge_index[0]), half 1 onto its second; each cell then takes (s[v0] + s[v1] + s[v2]) / 3.
    """
    idx = torch.cat([v_edge_index[0], v_edge_index[1]], dim=0)
    fwd, rev = torch.chunk(e, 2, dim=-1)
    vsum = scatter_add(torch.cat([fwd, rev], dim=0), idx, n_vertices)
    agg = (vsum.index_select(0, v_face[0]) + vsum.index_select(0, v_face[1])
           + vsum.index_select(0, v_face[2])) / 3.0
    return agg, vsum


def cell_block_two_hop(sd, prefix, x, e, v_edge_index, v_face, n_vertices):
    agg, _ = two_hop_aggregate(e, v_edge_index, v_face, n_vertices)
    return mlp_from_state(sd, prefix, torch.cat([x, agg], dim=-1))


def face_block_concat(sd, prefix, x, e, c_edge_index):
    """Face_Block, concat form (Fvgn.py:292-296 == Mgn.py:234-238): order is [e, x[row], x[col]]."""
    row, col = c_edge_index[0], c_edge_index[1]
    return mlp_from_state(sd, prefix, torch.cat([e, x[row], x[col]], dim=1))


def face_block_sum(sd, prefix, x, e, c_edge_index, e_asym=None):
    """Face_Block, sum form (Conservative.py:228-234); optional multiply by the asym encoding."""
    row, col = c_edge_index[0], c_edge_index[1]
    out = mlp_from_state(sd, prefix, torch.cat([e, x[row] + x[col]], dim=1))
    if e_asym is not None:
        out = out * e_asym
    return out


def cell_block_signed(sd, prefix, x, e, c_edge_index):
    """Cell_Block, signed direct edge->cell sum (Conservative.py:243-254):
    agg[c] = sum_{col(k)=c} e_k - sum_{row(k)=c} e_k."""
    row, col = c_edge_index[0], c_edge_index[1]
    idx = torch.cat([col, row], dim=0)
    agg = scatter_add(torch.cat([e, -e], dim=0), idx, x.shape[0])
    return mlp_from_state(sd, prefix, torch.cat([x, agg], dim=-1))


def cell_block_sym_asym_halves(sd, prefix, x, e, c_edge_index):
    """ConservativeE Cell_Block (Conservative.py:708-732): the first half of the edge latent is summed onto both
    cells of the face with the same sign, the second half with opposite signs; agg = cat[sym, asym]."""
    row, col = c_edge_index[0], c_edge_index[1]
    idx = torch.cat([col, row], dim=0)
    es, ea = torch.chunk(e, 2, dim=-1)
    sym = scatter_add(torch.cat([es, es], dim=0), idx, x.shape[0])
    asym = scatter_add(torch.cat([ea, -ea], dim=0), idx, x.shape[0])
    return mlp_from_state(sd, prefix, torch.cat([x, sym, asym], dim=-1))


def cell_block_hybrid(sd, prefix, x, e, c_edge_index, v_edge_index, v_face, n_vertices):
    """ConservativeF Cell_Block (Conservative.py:782-809): symmetric half via the vertices (the SAME half onto both
    vertices of a face, then the 3-vertex mean), antisymmetric half as a signed direct edge->cell sum."""
    es, ea = torch.chunk(e, 2, dim=-1)
    vidx = torch.cat([v_edge_index[0], v_edge_index[1]], dim=0)
    vsum = scatter_add(torch.cat([es, es], dim=0), vidx, n_vertices)
    cell_agg = (vsum.index_select(0, v_face[0]) + vsum.index_select(0, v_face[1]) + vsum.index_select(0, v_face[2])) / 3.0
    row, col = c_edge_index[0], c_edge_index[1]
    asym = scatter_add(torch.cat([ea, -ea], dim=0), torch.cat([col, row], dim=0), x.shape[0])
    return mlp_from_state(sd, prefix, torch.cat([x, cell_agg, asym], dim=-1))


def gn_block_dual(sd, i, x, e_s, e_a, c_edge_index):
    """ConservativeD GN_Block (Conservative.py:572-645): two edge streams.  Symmetric face block
    MLP(cat[e_s, x[row] + x[col]]) (SiLU + LayerNorm), antisymmetric face block antisym_MLP(cat[e_a, x[row] - x[col]])
    (bias-free tanh, no LayerNorm); the cell block sums the raw symmetric output onto both cells with equal signs and
    the raw antisymmetric output with opposite signs; residuals on x, e_s, e_a after the block."""
    p = f"processer_list.{i}"
    row, col = c_edge_index[0], c_edge_index[1]
    sr = mlp_from_state(sd, f"{p}.face_block_symm.face_mlp", torch.cat([e_s, x[row] + x[col]], dim=1))
    ar = mlp_from_state(sd, f"{p}.face_block_asym.face_mlp", torch.cat([e_a, x[row] - x[col]], dim=1), act="tanh")
    idx = torch.cat([col, row], dim=0)
    sym = scatter_add(torch.cat([sr, sr], dim=0), idx, x.shape[0])
    asym = scatter_add(torch.cat([ar, -ar], dim=0), idx, x.shape[0])
    xr = mlp_from_state(sd, f"{p}.cell_block.cell_mlp", torch.cat([x, sym, asym], dim=-1))
    return x + xr, e_s + sr, e_a + ar


def gn_block_dual_two_hop(sd, i, x, e_s, e_a, topo):
    """ConservativeH / J GN_Block (Conservative.py:1098-1184): cell block FIRST - the full symmetric stream summed
    onto both vertices of each face then the 3-vertex mean, the antisymmetric stream as a signed direct edge->cell
    sum, cell MLP on cat[x, sym, asym] - then the symmetric (x'[row] + x'[col]) and antisymmetric (x'[row] - x'[col])
    face blocks on the RAW cell output; residuals on all three streams."""
    p = f"processer_list.{i}"
    row, col = topo["c_edge_index"][0], topo["c_edge_index"][1]
    vidx = torch.cat([topo["v_edge_index"][0], topo["v_edge_index"][1]], dim=0)
    vsum = scatter_add(torch.cat([e_s, e_s], dim=0), vidx, topo["n_vertices"])
    vf = topo["v_face"]
    cell_agg = (vsum.index_select(0, vf[0]) + vsum.index_select(0, vf[1]) + vsum.index_select(0, vf[2])) / 3.0
    asym = scatter_add(torch.cat([e_a, -e_a], dim=0), torch.cat([col, row], dim=0), x.shape[0])
    xr = mlp_from_state(sd, f"{p}.cell_block.cell_mlp", torch.cat([x, cell_agg, asym], dim=-1))
    sr = mlp_from_state(sd, f"{p}.face_block_symm.face_mlp", torch.cat([e_s, xr[row] + xr[col]], dim=1))
    ar = mlp_from_state(sd, f"{p}.face_block_asym.face_mlp", torch.cat([e_a, xr[row] - xr[col]], dim=1), act="tanh")
    return x + xr, e_s + sr, e_a + ar


def decoder_even_odd(sd, e_s, e_a):
    """ConservativeH Decoder (Conservative.py:1186-1208): even head on cat[h+, h-^2] -> (u, v, p, |q|(2)); odd
    (antisymmetric) head on cat[h-, h+] -> sign in (-1, 1);  q_n = softplus(|q|) * tanh(odd)."""
    even = mlp_from_state(sd, "decoder.even_mlp", torch.cat([e_s, e_a ** 2], dim=-1))
    odd = torch.tanh(mlp_from_state(sd, "decoder.odd_mlp", torch.cat([e_a, e_s], dim=-1), act="tanh"))
    return torch.cat([even[:, 0:3], torch.nn.functional.softplus(even[:, 3:5]) * odd], dim=-1)


def vertex_block(e, v_edge_index, n_rows):
    """Vertex_Block (VertPot.py:217-222): full-width edge->vertex sum; the output has
    ``cell_graph.x.size(0)`` rows (N, not V) - rows >= V stay zero."""
    idx = torch.cat([v_edge_index[0], v_edge_index[1]], dim=0)
    return scatter_add(e.repeat(2, 1), idx, n_rows)


# --- encoder / block / decoder per family ---------------------------------------------------

def encoder_fwd(family, sd, c_x, f_x, f_x_asym=None):
    """Encoder.forward (Fvgn.py:257-266, Mgn.py:199-208, Conservative.py:191-202).
    Returns (x0, e0, e0_asym-or-None)."""
    if family in ("cons_a", "cons_b"):
        e0 = mlp_from_state(sd, "encoder.faceS_mlp", f_x)
        ea = mlp_from_state(sd, "encoder.faceA_mlp", f_x_asym, act="tanh")
        x0 = mlp_from_state(sd, "encoder.cell_mlp", c_x)
        return x0, e0, ea
    e0 = mlp_from_state(sd, "encoder.face_mlp", f_x)
    x0 = mlp_from_state(sd, "encoder.cell_mlp", c_x)
    return x0, e0, None


def gn_block_fwd(family, sd, i, x, e, topo, e_asym=None, bc_mask=None):
    """One GN_Block.  ``topo`` has c_edge_index, v_edge_index, v_face, n_vertices.
    The second sub-block consumes the first one's RAW output, the residuals are added after both
    (Fvgn.py:274-284, Mgn.py:216-226, Conservative.py:210-220, VertPot.py:195-210).
    Returns (x_new, e_new, vertex_x-or-None)."""
    p = f"processer_list.{i}"
    if family == "fvgn":
        xr = cell_block_two_hop(sd, f"{p}.cell_block.cell_mlp", x, e, topo["v_edge_index"],
                                topo["v_face"], topo["n_vertices"])
        er = face_block_concat(sd, f"{p}.face_block.face_mlp", xr, e, topo["c_edge_index"])
        return x + xr, e + er, None
    if family == "mgn":
        er = face_block_concat(sd, f"{p}.face_block.face_mlp", x, e, topo["c_edge_index"])
        xr = cell_block_two_hop(sd, f"{p}.cell_block.cell_mlp", x, er, topo["v_edge_index"],
                                topo["v_face"], topo["n_vertices"])
        return x + xr, e + er, None
    if family in ("cons_a", "cons_b"):      # ConservativeB = ConservativeA's encoder and blocks (Conservative.py:271-275)
        er = face_block_sum(sd, f"{p}.face_block.face_mlp", x, e, topo["c_edge_index"], e_asym)
        xr = cell_block_signed(sd, f"{p}.cell_block.cell_mlp", x, er, topo["c_edge_index"])
        return x + xr, e + er, None
    if family == "cons_e":      # face block (sum form) -> cell block on the RAW face output (Conservative.py:677-687)
        er = face_block_sum(sd, f"{p}.face_block.face_mlp", x, e, topo["c_edge_index"])
        xr = cell_block_sym_asym_halves(sd, f"{p}.cell_block.cell_mlp", x, er, topo["c_edge_index"])
        return x + xr, e + er, None
    if family == "cons_f":      # cell block (hybrid) -> face block (concat form) on the RAW cell output (:763-773)
        xr = cell_block_hybrid(sd, f"{p}.cell_block.cell_mlp", x, e, topo["c_edge_index"], topo["v_edge_index"],
                               topo["v_face"], topo["n_vertices"])
        er = face_block_concat(sd, f"{p}.face_block.face_mlp", xr, e, topo["c_edge_index"])
        return x + xr, e + er, None
    if family in ("cons_g", "cons_i"):
        # ConservativeG (Conservative.py:834-896): hybrid cell block (as F) -> SUM-form face block on the raw cell
        # output.  ConservativeI (:1250-1269) additionally keeps the previous latent on INFLOW / WALL faces.
        xr = cell_block_hybrid(sd, f"{p}.cell_block.cell_mlp", x, e, topo["c_edge_index"], topo["v_edge_index"],
                               topo["v_face"], topo["n_vertices"])
        er = face_block_sum(sd, f"{p}.face_block.face_mlp", xr, e, topo["c_edge_index"])
        e_new = e + er
        if family == "cons_i":
            e_new = e_new.clone()
            e_new[bc_mask] = e[bc_mask]
        return x + xr, e_new, None
    if family == "vertpot":
        xr = cell_block_two_hop(sd, f"{p}.node_block.cell_mlp", x, e, topo["v_edge_index"],
                                topo["v_face"], topo["n_vertices"])
        er = face_block_concat(sd, f"{p}.edge_block.face_mlp", xr, e, topo["c_edge_index"])
        # Vertex_Block runs on the face block's RAW output (c_graph at that point carries e')
        vx = vertex_block(er, topo["v_edge_index"], x.shape[0])
        return x + xr, e + er, vx
    raise ValueError(family)


def decoder_fwd(family, sd, x, e, vx=None):
    """Decoder.forward: edge head (Fvgn.py:327-333), node head (Mgn.py:269-275),
    edge+vertex heads (VertPot.py:224-231)."""
    if family == "mgn":
        return mlp_from_state(sd, "decoder.face_mlp", x)
    if family == "cons_b":      # node head of ConservativeB (Conservative.py:406-414)
        return mlp_from_state(sd, "decoder.node_mlp", x)
    if family == "vertpot":
        return (mlp_from_state(sd, "decoder.edge_mlp", e), mlp_from_state(sd, "decoder.vertex_mlp", vx))
    return mlp_from_state(sd, "decoder.face_mlp", e)


def processor_fwd(family, sd, c_x, f_x, topo, mp_num, f_x_asym=None, keep_blocks=False, bc_mask=None):
    """encoder -> mp_num x GN_Block -> decoder on already-normalised inputs.

    ConservativeA quirk reproduced: GN_Block returns a fresh Data without ``edge_attr_asym`` so
    the asym multiply fires in block 0 only (Conservative.py:220, 232-233)."""
    if family == "fvgn_f":
        # FvgnF (Fvgn.py:881-1002): ONE shared GN_Block (keys gn_block.*) applied mp_num times, every MLP input gets
        # the constant column (step + 1) / mp_num appended
        x, e, _ = encoder_fwd("fvgn", sd, c_x, f_x)
        out = {"x0": x, "e0": e}
        per_block = []
        row, col = topo["c_edge_index"][0], topo["c_edge_index"][1]
        for i in range(mp_num):
            step = (i + 1) / mp_num
            agg, _ = two_hop_aggregate(e, topo["v_edge_index"], topo["v_face"], topo["n_vertices"])
            xr = mlp_from_state(sd, "gn_block.cell_block.cell_mlp",
                                torch.cat([x, agg, torch.full((x.shape[0], 1), step, dtype=x.dtype)], dim=-1))
            er = mlp_from_state(sd, "gn_block.face_block.face_mlp",
                                torch.cat([e, xr[row], xr[col], torch.full((e.shape[0], 1), step, dtype=e.dtype)], dim=1))
            x, e = x + xr, e + er
            if keep_blocks:
                per_block.append((x, e))
        out.update({"x": x, "e": e, "vx": None, "blocks": per_block, "dec": mlp_from_state(sd, "decoder.face_mlp", e)})
        return out
    if family == "cons_h":
        x, e, ea = encoder_fwd("cons_a", sd, c_x, f_x, f_x_asym)
        out = {"x0": x, "e0": e, "e0_asym": ea}
        per_block = []
        for i in range(mp_num):
            x, e, ea = gn_block_dual_two_hop(sd, i, x, e, ea, topo)
            if keep_blocks:
                per_block.append((x, e, ea))
        out.update({"x": x, "e": e, "ea": ea, "vx": None, "blocks": per_block, "dec": decoder_even_odd(sd, e, ea)})
        return out
    if family == "cons_d":
        x, e, ea = encoder_fwd("cons_a", sd, c_x, f_x, f_x_asym)      # same encoder containers as ConservativeA
        out = {"x0": x, "e0": e, "e0_asym": ea}
        per_block = []
        for i in range(mp_num):
            x, e, ea = gn_block_dual(sd, i, x, e, ea, topo["c_edge_index"])
            if keep_blocks:
                per_block.append((x, e, ea))
        out.update({"x": x, "e": e, "ea": ea, "vx": None, "blocks": per_block})
        # Decoder (Conservative.py:647-658): final_mlp(symm_mlp(e_s) + asym_mlp(e_a))
        comb = mlp_from_state(sd, "decoder.symm_mlp", e) + mlp_from_state(sd, "decoder.asym_mlp", ea, act="tanh")
        out["dec"] = mlp_from_state(sd, "decoder.final_mlp", comb, act="tanh")
        return out
    x, e, ea = encoder_fwd(family, sd, c_x, f_x, f_x_asym)
    out = {"x0": x, "e0": e}
    vx = None
    per_block = []
    for i in range(mp_num):
        x, e, vx = gn_block_fwd(family, sd, i, x, e, topo, e_asym=ea if i == 0 else None, bc_mask=bc_mask)
        if keep_blocks:
            per_block.append((x, e))
    out.update({"x": x, "e": e, "vx": vx, "blocks": per_block})
    out["dec"] = decoder_fwd(family, sd, x, e, vx)
    return out
