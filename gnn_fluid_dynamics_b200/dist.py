"""Multi-GPU execution of the hot path.

* Independent meshes (BASELINE.json configs 2, 5): plain data parallelism - every rank runs the whole
  path on its own meshes, no data-path collective; training adds one gradient all-reduce per step
  (``allreduce_gradients``).
* One large mesh (config 4): domain decomposition (``partition.py``) with ONE halo exchange of ghost-cell
  latents per GN_Block.  The send side packs boundary rows with ``gnnfd_gather_rows``; the transfer is NCCL
  point-to-point over NVLink (``torch.distributed.batch_isend_irecv``); the receive side lands directly in the
  ghost rows (contiguous per owner rank, no unpack).  Face latents are never communicated.

``InProcessTransport`` runs all partitions in one process on one GPU (same kernels, same plan, copies
instead of NCCL): it is how the single-GPU tests pin the partitioned result bit-for-bit to the unpartitioned
one; ``TorchDistTransport`` is the one-process-per-GPU transport.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch

from . import ops
from .ops import Seg
from .partition import Partition
from .processor import H, weights_of
from ._lib import SEG_GATHER, SEG_MEAN3


@dataclass
class PartState:
    """Device-side state of one partition: plan, local topology, latents (rows: owned cells, then ghosts)."""
    part: Partition
    topo: object                     # MeshTopology of the local sub-mesh
    x: Optional[torch.Tensor] = None
    e: Optional[torch.Tensor] = None
    send_idx: Optional[dict] = None  # peer -> int32 device tensor
    xs: Optional[torch.Tensor] = None    # split precisions: 16-bit hi|lo shadow of the gathered cell matrix (processor.Fast)
    e_asym: Optional[torch.Tensor] = None    # 'cons_a': the antisymmetric face encoding multiplied into block 0's face output

    def device_plan(self, device):
        if self.send_idx is None:
            self.send_idx = {p: idx.to(torch.int32).to(device) for p, idx in self.part.send.items()}
        return self


def _pack_cuda(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return ops.gather_rows(t, idx)


class InProcessTransport:
    """All partitions live in this process: the exchange is a row copy between their tensors."""

    def __init__(self, pack: Callable = _pack_cuda):
        self.pack = pack
        self.bytes_sent = 0

    def exchange(self, states: Sequence[PartState], get: Callable[[PartState], torch.Tensor]):
        by_rank = {s.part.rank: s for s in states}
        for dst in states:
            t = get(dst)
            for peer, (start, cnt) in dst.part.recv.items():
                src = by_rank[peer]
                src.device_plan(get(src).device)
                rows = self.pack(get(src), src.send_idx[dst.part.rank])
                t[start:start + cnt].copy_(rows)
                self.bytes_sent += rows.numel() * 4


class TorchDistTransport:
    """One partition per process (torch.distributed, NCCL on GPUs / gloo in the CPU tests): grouped
    point-to-point sends and receives, all posted together so NCCL runs them as one group."""

    def __init__(self, group=None, pack: Callable = _pack_cuda):
        import torch.distributed as dist
        self.dist, self.group, self.pack = dist, group, pack
        self.bytes_sent = 0

    def exchange(self, states: Sequence[PartState], get: Callable[[PartState], torch.Tensor]):
        dist = self.dist
        (s,) = states
        t = get(s)
        s.device_plan(t.device)
        ops_, keep = [], []
        for peer, (start, cnt) in sorted(s.part.recv.items()):
            ops_.append(dist.P2POp(dist.irecv, t[start:start + cnt], peer, self.group))
        for peer, idx in sorted(s.send_idx.items()):
            buf = self.pack(t, idx)
            keep.append(buf)
            self.bytes_sent += buf.numel() * 4
            ops_.append(dist.P2POp(dist.isend, buf, peer, self.group))
        if ops_:
            for req in dist.batch_isend_irecv(ops_):
                req.wait()


def _edge_segs(e, xs, topo):
    return [Seg(e), Seg(xs, SEG_GATHER, (topo.row,)), Seg(xs, SEG_GATHER, (topo.col,))]


def run_processor_partitioned(family: str, blocks, states: List[PartState], transport, prec: int):
    """The GN_Blocks over partitioned latents.  Edge phases run on all local faces, node phases on owned cells
    only; exactly one halo exchange per block (MGN: block-input x, skipped for block 0 where the ghosts' x is the
    locally encoded one; FVGN: the raw cell-MLP output x')."""
    if family == "cons_a":
        return _run_processor_partitioned_cons_a(blocks, states, transport, prec)
    if family not in ("mgn", "fvgn"):
        raise NotImplementedError(f"domain decomposition covers the 'mgn', 'fvgn' and 'cons_a' data-flows, not {family!r}")
    if ops.split_dtype(prec) is not None and all(s.xs is not None for s in states):
        return _run_processor_partitioned_fast(family, blocks, states, transport, prec)
    for i, blk in enumerate(blocks):
        we, wn = weights_of(blk.face_block.face_mlp), weights_of(blk.cell_block.cell_mlp)
        if family == "mgn":
            if i > 0:
                transport.exchange(states, lambda s: s.x)
            for s in states:
                topo, n_own = s.topo, s.part.n_owned
                e_raw, e_new = ops.mlp_forward(_edge_segs(s.e, s.x, topo), we, s.e.shape[0], prec, residual=s.e,
                                               want_raw=True, want_sum=True)
                vsum = ops.segment_sum(e_raw, e_raw, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm,
                                       topo.n_vertices)
                x_new = torch.empty_like(s.x)
                ops.mlp_forward([Seg(s.x), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=s.x,
                                want_raw=False, want_sum=True, out_sum=x_new)
                s.x, s.e = x_new, e_new
        else:
            for s in states:
                topo, n_own = s.topo, s.part.n_owned
                vsum = ops.segment_sum(s.e, s.e, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm,
                                       topo.n_vertices)
                s.x_raw, x_new = torch.empty_like(s.x), torch.empty_like(s.x)
                ops.mlp_forward([Seg(s.x), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=s.x,
                                want_raw=True, want_sum=True, out_raw=s.x_raw, out_sum=x_new)
                s.x = x_new
            transport.exchange(states, lambda s: s.x_raw)
            for s in states:
                _, s.e = ops.mlp_forward(_edge_segs(s.e, s.x_raw, s.topo), we, s.e.shape[0], prec, residual=s.e,
                                         want_raw=False, want_sum=True)
    return states


def _run_processor_partitioned_cons_a(blocks, states: List[PartState], transport, prec: int):
    """ConservativeA's GN_Blocks (Conservative.py:210-254) over partitioned latents: face block on all local faces
    (``x[row] + x[col]`` gathers ghost cells), signed edge->cell sum and cell block on the owned cells, ONE exchange of the
    block-input ``x`` per block (skipped for block 0: ghosts are encoded locally).  The vertex-star halo of the plan is a
    superset of the face-neighbour ring this data-flow needs (every face of an owned cell is local), and local faces keep
    global order, so the owned rows are bit-identical to the single-GPU result."""
    from ._lib import SEG_SUM2
    from .processor import fast_mode
    # inference: the residual streams advance in place through the TMA-store epilogue, exactly as processor.run_processor
    # does for this family (the same kernels on the same rows: bit-identical owned rows either way)
    inplace = fast_mode(blocks, [t for s in states for t in (s.x, s.e)], prec)
    for i, blk in enumerate(blocks):
        we, wn = weights_of(blk.face_block.face_mlp), weights_of(blk.cell_block.cell_mlp)
        if i > 0:
            transport.exchange(states, lambda s: s.x)
        for s in states:
            topo, n_own = s.topo, s.part.n_owned
            segs = [Seg(s.e), Seg(s.x, SEG_SUM2, (topo.row, topo.col))]
            e_raw, e_new = ops.mlp_forward(segs, we, s.e.shape[0], prec, mul=s.e_asym if i == 0 else None, residual=s.e,
                                           want_raw=True, want_sum=True, out_sum=s.e if inplace else None)
            off, perm = topo.build_cell_csr()
            agg = ops.segment_sum(e_raw, e_raw, 0, 0, H, -1.0, off, perm, n_own)
            x_new = s.x if inplace else torch.empty_like(s.x)     # ghost rows: filled by the next block's exchange
            ops.mlp_forward([Seg(s.x), Seg(agg)], wn, n_own, prec, residual=s.x, want_raw=False, want_sum=True, out_sum=x_new)
            s.x, s.e = x_new, e_new
    return states


def _shadow_rows(s: PartState) -> torch.Tensor:
    """The split shadow viewed as fp32 rows (512 B per cell, like a latent row): what the halo exchange moves."""
    return s.xs.view(torch.float32)


def _run_processor_partitioned_fast(family: str, blocks, states: List[PartState], transport, prec: int):
    """Same data-flow with the single-GPU inference fast path's kernels (``processor.Fast``): x / e updated in place,
    the gathered cell matrix handed over as its 16-bit split shadow ``s.xs`` (TMA gather4 in the edge kernel) - and it
    is the SHADOW's ghost rows that travel in the halo exchange (512 B per ghost cell, the size of an fp32 row), so the
    fp32 ghost rows are never needed.  Bit-identical to ``processor.run_processor`` in fast mode."""
    for i, blk in enumerate(blocks):
        we, wn = weights_of(blk.face_block.face_mlp), weights_of(blk.cell_block.cell_mlp)

        def gsegs(s):
            xf = _shadow_rows(s)
            return [Seg(s.e), Seg(xf, SEG_GATHER, (s.topo.row,), split=s.xs), Seg(xf, SEG_GATHER, (s.topo.col,), split=s.xs)]

        if family == "mgn":
            if i > 0:
                transport.exchange(states, _shadow_rows)
            for s in states:
                topo, n_own = s.topo, s.part.n_owned
                e_raw, _ = ops.mlp_forward(gsegs(s), we, s.e.shape[0], prec, residual=s.e, want_raw=True, want_sum=True,
                                           out_sum=s.e)
                vsum = ops.segment_sum(e_raw, e_raw, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm,
                                       topo.n_vertices)
                ops.mlp_forward([Seg(s.x), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=s.x,
                                want_raw=False, want_sum=True, out_sum=s.x, out_split=s.xs[:n_own], split_of_sum=True)
        else:
            for s in states:
                topo, n_own = s.topo, s.part.n_owned
                vsum = ops.segment_sum(s.e, s.e, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm,
                                       topo.n_vertices)
                ops.mlp_forward([Seg(s.x), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=s.x,
                                want_raw=False, want_sum=True, out_sum=s.x, out_split=s.xs[:n_own])
            transport.exchange(states, _shadow_rows)
            for s in states:
                ops.mlp_forward(gsegs(s), we, s.e.shape[0], prec, residual=s.e, want_raw=False, want_sum=True, out_sum=s.e)
    return states


def encode_process_decode_partitioned(model, states: List[PartState], inputs, transport):
    """encoder -> partitioned GN_Blocks -> decoder for an Fvgn/Mgn-family model.  ``inputs[k]`` = (c_x, f_x) of
    partition k (normalised, local rows).  Returns per partition (x[n_owned], e[E_loc], decoder output): the node
    decoder (MGN) covers the owned cells, the edge decoder (FVGN) all local faces."""
    from . import processor as P
    prec = model.prec
    if model.family == "cons_a":       # inputs[k] = (c_x, f_x_symm, f_x_asym); edge decoder (Conservative.py:164-189)
        from ._lib import ACT_TANH
        for s, (c_x, f_s, f_a) in zip(states, inputs):
            s.e = P.mlp_rows(model.encoder.faceS_mlp, f_s, prec)
            s.e_asym = P.mlp_rows(model.encoder.faceA_mlp, f_a, prec, act=ACT_TANH)
            s.x = P.mlp_rows(model.encoder.cell_mlp, c_x, prec)
            s.xs = None
        run_processor_partitioned("cons_a", model.processer_list, states, transport, prec)
        return [(s.x[:s.part.n_owned], s.e, P.mlp_rows(model.decoder.face_mlp, s.e, prec)) for s in states]
    sdt = ops.split_dtype(prec)
    for s, (c_x, f_x) in zip(states, inputs):
        s.e = P.mlp_rows(model.encoder.face_mlp, f_x, prec)
        if sdt is not None:      # inference fast path (see _run_processor_partitioned_fast)
            if s.xs is None or s.xs.shape[0] != c_x.shape[0] or s.xs.dtype != sdt:
                s.xs = torch.empty(c_x.shape[0], 2 * H, dtype=sdt, device=c_x.device)
            s.x, _ = ops.mlp_forward([Seg(c_x.contiguous())], weights_of(model.encoder.cell_mlp), c_x.shape[0], prec,
                                     out_split=s.xs if model.family == "mgn" else None)
        else:
            s.xs = None
            s.x = P.mlp_rows(model.encoder.cell_mlp, c_x, prec)     # ghosts encoded locally: no exchange before block 0
    run_processor_partitioned(model.family, model.processer_list, states, transport, prec)
    outs = []
    for s in states:
        n_own = s.part.n_owned
        if model.family == "mgn":
            dec = P.mlp_rows(model.decoder.face_mlp, s.x[:n_own], prec)
        else:
            dec = P.mlp_rows(model.decoder.face_mlp, s.e, prec)
        outs.append((s.x[:n_own], s.e, dec))
    return outs


def allreduce_gradients(params, world: int, group=None):
    """Data-parallel training over independent meshes: ONE flat all-reduce (average) of every gradient per step (the
    reference's DDP wrapper is bypassed by its own trainer, src/train.py:165; this is the working equivalent).
    Three device operations regardless of the parameter count: one multi-tensor pack, one NCCL all-reduce with the
    averaging done by the collective (ReduceOp.AVG; gloo: sum then one scale), one multi-tensor unpack."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if world <= 1 or not grads:
        return
    flat = torch.empty(sum(g.numel() for g in grads), dtype=grads[0].dtype, device=grads[0].device)
    views = list(flat.split([g.numel() for g in grads]))
    torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, group=group)
        flat /= world
    torch._foreach_copy_([g.view(-1) if g.is_contiguous() else g for g in grads],
                         [v if g.is_contiguous() else v.view_as(g) for v, g in zip(views, grads)])


class PartitionedRollout:
    """Autoregressive rollout of ONE large mesh over several partitions (BASELINE.json config 4): per step the
    encoder / decoder / integrator run locally, the GN_Blocks exchange ghost latents once per block, and the state
    advance exchanges the ghost cells' new velocity (2 floats per ghost cell) once per step.

    ``local_graphs[k]`` is partition k's [c_graph, f_graph, v_graph] on its device (``partition.local_graphs``);
    under torch.distributed each process holds exactly one partition."""

    def __init__(self, model, parts: Sequence[Partition], local_graphs: Sequence[list], transport, peer=None):
        """``peer`` = (PeerBuffers, row_enc, col_enc) switches the per-block halo from NCCL send/receive to
        peer-memory gathers inside the fused edge kernel (one partition per process only)."""
        from .graph import Data
        from .topology import MeshTopology
        if model.family not in ("mgn", "fvgn"):
            raise NotImplementedError("partitioned rollout covers the Mgn / Fvgn families")
        self.model, self.transport = model.eval(), transport
        self.graphs = list(local_graphs)
        self.states = []
        for p, g in zip(parts, self.graphs):
            topo = MeshTopology.from_graphs(g).validate()
            self.states.append(PartState(part=p, topo=topo).device_plan(g[0].x.device))
        self._Data = Data
        self.peer = peer
        if peer is not None and len(self.states) != 1:
            raise RuntimeError("peer-memory halo runs one partition per process")

    @torch.no_grad()
    def step(self):
        """One timestep on every held partition; returns the owned cells' new velocity per partition."""
        model = self.model
        norm, inputs = [], []
        for g in self.graphs:
            gn = model.normalizer.input([t.clone() for t in g])
            norm.append(gn)
            inputs.append((gn[0].x, gn[1].x))
        if self.peer is None:
            outs = encode_process_decode_partitioned(model, self.states, inputs, self.transport)
        else:
            from . import processor as P
            bufs, row_enc, col_enc = self.peer
            (s,), ((c_x, f_x),) = self.states, inputs
            n_own = s.part.n_owned
            e0 = P.mlp_rows(model.encoder.face_mlp, f_x, model.prec)
            x0 = P.mlp_rows(model.encoder.cell_mlp, c_x[:n_own], model.prec)      # ghosts are never materialised
            x, e = run_processor_peer(model.family, model.processer_list, s, x0, e0, bufs, row_enc, col_enc, model.prec)
            dec = P.mlp_rows(model.decoder.face_mlp, x if model.family == "mgn" else e, model.prec)
            outs = [(x, e, dec)]
        vels = []
        for s, g, gn, (_, _, dec) in zip(self.states, self.graphs, norm, outs):
            n_own = s.part.n_owned
            c, f, _ = gn
            if model.family == "mgn":
                output = [dec, None, None]
            else:
                # the local topology rides along so the integrator runs the same fused finite-volume kernels as the
                # single-GPU model (bit-identical velocities on the owned cells)
                c_view = self._Data(x=c.x, normal=c.normal[:n_own], volume=c.volume, edge_index=c.edge_index, dt=c.dt,
                                    topology=s.topo)
                output = [model.integrator(dec, c_view, f, c.dt), dec, None]
            output = model.normalizer.output(output, inverse=True)
            vel = g[0].x[:n_own, :2] + output[0][:, 0:2]
            g[0].x[:n_own, :2] = vel
            vels.append(vel)
        by_id = {id(st): g for st, g in zip(self.states, self.graphs)}
        self.transport.exchange(self.states, lambda s: by_id[id(s)][0].x)
        for g in self.graphs:
            c, f, _ = g
            u = c.x[:, :2]
            dv = u[c.edge_index[0]] - u[c.edge_index[1]]
            if model.family == "mgn":
                mask = f.boundary_mask
            else:
                mask = ((f.type == 2) | (f.type == 1)).squeeze(-1)
            f.x[:, 0:2] = torch.where(mask.unsqueeze(-1), f.y[:, 0:2], dv)
        return vels


# ------------------------------------------------------------------------------------------------------------
# Peer-memory halo: no exchange step at all.  The cell latents of every rank live in buffers that all ranks map
# (CUDA IPC, NVLink P2P); the fused edge kernel gathers x[row] / x[col] of ghost cells straight from the owning
# GPU's HBM while it computes (gnnfd_mlp_args.peer_base / peer_shift), tile by tile, so the transfer overlaps
# the MMA work and ghost rows are never materialised locally.  One cross-rank barrier per GN_Block orders
# "owner wrote block i's rows" before "peers read them"; the latents are double-buffered so the next block's
# writes cannot race the previous block's remote reads.

PEER_SHIFT = 28          # gather index = (owner rank << 28) | row in the owner's buffer


class PeerBuffers:
    """``n_buffers`` matrices [n_rows, width] per rank, each mapped into every process of the group."""

    def __init__(self, n_rows: int, width: int, device, world: int, rank: int, n_buffers: int = 2, group=None):
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        from ._lib import check, lib
        self.world, self.rank = world, rank
        self.local = [torch.zeros(max(n_rows, 1), width, dtype=torch.float32, device=device) for _ in range(n_buffers)]
        payload = [reduce_tensor(t) for t in self.local]                  # CUDA IPC handles (picklable)
        gathered = [None] * world
        dist.all_gather_object(gathered, payload, group=group)
        self.views = []                                                    # views[buffer][rank] -> tensor
        for b in range(n_buffers):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(self.local[b])
                else:
                    fn, fargs = gathered[r][b]
                    fargs = list(fargs)
                    # torch's CUDA-IPC rebuild tuple is not a public contract: refuse to guess if its layout moved
                    if getattr(fn, "__name__", "") != "rebuild_cuda_tensor" or len(fargs) < 8 or not isinstance(fargs[6], int):
                        raise RuntimeError("PeerBuffers: torch.multiprocessing.reductions.reduce_tensor returned an unexpected "
                                           f"rebuild recipe ({getattr(fn, '__name__', fn)}, {len(fargs)} arguments); this torch "
                                           "version needs a new mapping of the owner-device argument")
                    owner_device = fargs[6]
                    # open the handle with THIS rank's device current (cudaIpcOpenMemHandle + lazy peer access maps the
                    # owner's memory for the opening device); the tensor is then "on" our device but lives in the
                    # owner's HBM and every access crosses NVLink
                    fargs[6] = device.index if device.index is not None else torch.cuda.current_device()
                    check(lib.gnnfd_enable_peer_access(owner_device), "gnnfd_enable_peer_access")
                    row.append(fn(*fargs))
            self.views.append(row)
        self._flag = torch.zeros(1, device=device)
        self._dist, self._group = dist, group
        dist.barrier(group=group)

    def barrier(self):
        """Stream-ordered cross-rank barrier: every rank's prior kernels are complete before any rank's later
        kernels start (a 1-element NCCL all-reduce on the current stream)."""
        self._dist.all_reduce(self._flag, group=self._group)


def peer_indices(part: Partition, parts_send: dict, device) -> tuple:
    """Peer-encoded copies of the partition's cell indices: local cell c < n_owned -> (rank << S) | c; ghost from
    peer a at position k of its receive range -> (a << S) | (row of that cell in a's buffer).  ``parts_send[a]`` is
    partition a's send list towards this rank (== the rows this rank's ghosts occupy in a's buffer)."""
    enc = torch.empty(part.n_local, dtype=torch.int64)
    enc[:part.n_owned] = (part.rank << PEER_SHIFT) | torch.arange(part.n_owned)
    for peer, (start, cnt) in part.recv.items():
        enc[start:start + cnt] = (peer << PEER_SHIFT) | parts_send[peer]
    row = enc[part.c_edge_index[0]].to(torch.int32).to(device)
    col = enc[part.c_edge_index[1]].to(torch.int32).to(device)
    return row, col


def run_processor_peer(family: str, blocks, state: PartState, x0_owned: torch.Tensor, e0: torch.Tensor,
                       bufs: PeerBuffers, row_enc: torch.Tensor, col_enc: torch.Tensor, prec: int):
    """GN_Blocks of one partition with peer-memory gathers (one process per GPU).  ``x0_owned`` = encoded owned
    cells, ``e0`` = encoded local faces.  Returns (x[n_owned], e[E_loc])."""
    if family not in ("mgn", "fvgn"):
        raise NotImplementedError(f"peer-memory halo covers the 'mgn' and 'fvgn' data-flows, not {family!r}")
    topo, n_own = state.topo, state.part.n_owned
    e = e0
    if family == "mgn":
        bufs.local[0][:n_own].copy_(x0_owned)
        bufs.barrier()
        for i, blk in enumerate(blocks):
            we, wn = weights_of(blk.face_block.face_mlp), weights_of(blk.cell_block.cell_mlp)
            cur, nxt = i % 2, (i + 1) % 2
            x_cur = bufs.local[cur]
            segs = [Seg(e), Seg(x_cur, SEG_GATHER, (row_enc,)), Seg(x_cur, SEG_GATHER, (col_enc,))]
            e_raw, e = ops.mlp_forward(segs, we, e.shape[0], prec, residual=e, want_raw=True, want_sum=True,
                                       peer=(bufs.views[cur], PEER_SHIFT))
            vsum = ops.segment_sum(e_raw, e_raw, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices)
            ops.mlp_forward([Seg(x_cur), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=x_cur,
                            want_raw=False, want_sum=True, out_sum=bufs.local[nxt])
            bufs.barrier()
        return bufs.local[len(blocks) % 2][:n_own], e
    x = x0_owned
    for i, blk in enumerate(blocks):
        we, wn = weights_of(blk.face_block.face_mlp), weights_of(blk.cell_block.cell_mlp)
        raw = bufs.local[i % 2]
        vsum = ops.segment_sum(e, e, 0, H // 2, H // 2, 1.0, topo.vtx_offsets, topo.vtx_perm, topo.n_vertices)
        _, x = ops.mlp_forward([Seg(x), Seg(vsum, SEG_MEAN3, topo.vf)], wn, n_own, prec, residual=x,
                               want_raw=True, want_sum=True, out_raw=raw)
        bufs.barrier()
        segs = [Seg(e), Seg(raw, SEG_GATHER, (row_enc,)), Seg(raw, SEG_GATHER, (col_enc,))]
        _, e = ops.mlp_forward(segs, we, e.shape[0], prec, residual=e, want_raw=False, want_sum=True,
                               peer=(bufs.views[i % 2], PEER_SHIFT))
    # the last block's remote reads of raw[(L - 1) % 2] must finish everywhere before a following call's block 0 (odd L:
    # the same buffer) overwrites it
    bufs.barrier()
    return x, e
