"""GPU: the domain-decomposed processor equals the unpartitioned one BIT FOR BIT on every owned cell and every
local face (row-independent kernels + order-preserving local CSRs), with all partitions emulated on one GPU
(InProcessTransport) and - when the box has >= 2 GPUs - with one process per GPU over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import build_model, golden_graphs

pytestmark = pytest.mark.gpu


def _setup(name, n_cells, world, device):
    from gnn_fluid_dynamics_b200.partition import local_graphs, partition_mesh
    model = build_model(name).eval()
    _, graphs = golden_graphs(name, n_cells=n_cells, mesh_seed=21, feat_seed=22)
    graphs = model.normalizer.input([g.clone() for g in graphs])
    c, f, v = graphs
    parts = partition_mesh(c.edge_index, v.edge_index, v.face, c.pos[:, 0], world, f_face=f.face)
    return model.to(device), graphs, parts, [local_graphs(graphs, p) for p in parts]


def _states(parts, locals_, device, only=None):
    from gnn_fluid_dynamics_b200.dist import PartState
    from gnn_fluid_dynamics_b200.topology import MeshTopology
    states, inputs = [], []
    for p, g in zip(parts, locals_):
        if only is not None and p.rank != only:
            continue
        gd = [t.to(device) for t in g]
        topo = MeshTopology.from_graphs(gd).validate()
        states.append(PartState(part=p, topo=topo))
        inputs.append((gd[0].x, gd[1].x))
    return states, inputs


@pytest.mark.parametrize("world", [2, 5])
def test_partitioned_conservative_equals_unpartitioned_bitwise(world):
    """ConservativeA (face-flux message passing: SUM2 face block, signed edge->cell sum, block-0 asym multiply) over a
    partitioned mesh == the single-GPU result, bit for bit, on owned cells / local faces (both advance the residual
    streams in place through the TMA-store epilogue)."""
    from gnn_fluid_dynamics_b200.dist import InProcessTransport, PartState, encode_process_decode_partitioned
    from gnn_fluid_dynamics_b200.topology import MeshTopology, get_topology
    dev = torch.device("cuda:0")
    model, graphs, parts, locals_ = _setup("ConservativeA", 4000, world, dev)
    gd = [g.to(dev) for g in graphs]
    with torch.no_grad():
        topo = get_topology(gd, need_cell_csr=True, two_hop=False).validate()
        x, e, dec = model.encode_process_decode(gd[0].x, gd[1].x_symm, gd[1].x_asym, topo)
        states, inputs = [], []
        for p, g in zip(parts, locals_):
            gl = [t.to(dev) for t in g]
            states.append(PartState(part=p, topo=MeshTopology.from_graphs(gl).validate()))
            inputs.append((gl[0].x, gl[1].x_symm, gl[1].x_asym))
        transport = InProcessTransport()
        outs = encode_process_decode_partitioned(model, states, inputs, transport)
    assert transport.bytes_sent > 0
    for p, (xp, ep, dp) in zip(parts, outs):
        cells, faces = p.cells[:p.n_owned].to(dev), p.faces.to(dev)
        assert torch.equal(xp, x[cells]) and torch.equal(ep, e[faces]) and torch.equal(dp, dec[faces])


@pytest.mark.parametrize("name", ["MgnA", "FvgnA"])
@pytest.mark.parametrize("world", [2, 5])
def test_partitioned_equals_unpartitioned_bitwise(name, world):
    from gnn_fluid_dynamics_b200.dist import InProcessTransport, encode_process_decode_partitioned
    from gnn_fluid_dynamics_b200.topology import get_topology
    dev = torch.device("cuda:0")
    model, graphs, parts, locals_ = _setup(name, 4000, world, dev)
    gd = [g.to(dev) for g in graphs]
    with torch.no_grad():
        x, e, dec = model.encode_process_decode(gd[0].x, gd[1].x, get_topology(gd).validate())
        states, inputs = _states(parts, locals_, dev)
        transport = InProcessTransport()
        outs = encode_process_decode_partitioned(model, states, inputs, transport)
    assert transport.bytes_sent > 0
    for p, (xp, ep, dp) in zip(parts, outs):
        cells, faces = p.cells[:p.n_owned].to(dev), p.faces.to(dev)
        assert torch.equal(xp, x[cells])
        assert torch.equal(ep, e[faces])
        assert torch.equal(dp, dec[cells] if name == "MgnA" else dec[faces])


def _nccl_rank(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from gnn_fluid_dynamics_b200.dist import TorchDistTransport, encode_process_decode_partitioned
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model, graphs, parts, locals_ = _setup(name, 4000, world, dev)
        states, inputs = _states(parts, locals_, dev, only=rank)
        with torch.no_grad():
            (xp, ep, dp), = encode_process_decode_partitioned(model, states, inputs, TorchDistTransport())
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=xp.cpu().numpy(), e=ep.cpu().numpy(), dec=dp.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("name", ["MgnA", "FvgnA"])
def test_nccl_halo_exchange_equals_single_gpu(tmp_path, name):
    from gnn_fluid_dynamics_b200.topology import get_topology
    world = min(torch.cuda.device_count(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_nccl_rank, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    model, graphs, parts, _ = _setup(name, 4000, world, dev)
    gd = [g.to(dev) for g in graphs]
    with torch.no_grad():
        x, e, dec = model.encode_process_decode(gd[0].x, gd[1].x, get_topology(gd).validate())
    for p in parts:
        d = np.load(tmp_path / f"rank{p.rank}.npz")
        cells, faces = p.cells[:p.n_owned], p.faces
        assert np.array_equal(d["x"], x.cpu()[cells].numpy())
        assert np.array_equal(d["e"], e.cpu()[faces].numpy())
        assert np.array_equal(d["dec"], (dec.cpu()[cells] if name == "MgnA" else dec.cpu()[faces]).numpy())


@pytest.mark.parametrize("name", ["MgnA", "FvgnA"])
def test_partitioned_rollout_equals_single_gpu_rollout(name):
    """5 autoregressive steps over 3 partitions (in-process transport) == the single-GPU RolloutEngine."""
    from gnn_fluid_dynamics_b200.dist import InProcessTransport, PartitionedRollout
    from gnn_fluid_dynamics_b200.partition import local_graphs, partition_mesh
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    dev = torch.device("cuda:0")
    model = build_model(name).to(dev).eval()
    _, graphs = golden_graphs(name, n_cells=3000, mesh_seed=41, feat_seed=42)
    c, f, v = graphs
    parts = partition_mesh(c.edge_index, v.edge_index, v.face, c.pos[:, 0], 3, f_face=f.face)
    pr = PartitionedRollout(model, parts, [[t.to(dev) for t in local_graphs(graphs, p)] for p in parts], InProcessTransport())
    # (the literal loop body: the partitioned step uses the same tensor glue, so the comparison stays bitwise)
    eng = RolloutEngine(model, [g.clone().to(dev) for g in graphs], cuda_graph=False, fused_step=False)
    for _ in range(5):
        vels = pr.step()
        ref = eng.step()
        for p, vp in zip(parts, vels):
            assert torch.equal(vp, ref[p.cells[:p.n_owned].to(dev)])


def _peer_rank(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from gnn_fluid_dynamics_b200 import processor as P
    from gnn_fluid_dynamics_b200.dist import PeerBuffers, peer_indices, run_processor_peer
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model, graphs, parts, locals_ = _setup(name, 4000, world, dev)
        part = parts[rank]
        states, inputs = _states(parts, locals_, dev, only=rank)
        (state,), ((c_x, f_x),) = states, inputs
        bufs = PeerBuffers(max(p.n_owned for p in parts), 128, dev, world, rank)
        row_enc, col_enc = peer_indices(part, {a: parts[a].send[rank] for a in part.recv}, dev)
        with torch.no_grad():
            e0 = P.mlp_rows(model.encoder.face_mlp, f_x, model.prec)
            x0 = P.mlp_rows(model.encoder.cell_mlp, c_x[:part.n_owned], model.prec)     # owned cells only: no ghosts anywhere
            x, e = run_processor_peer(model.family, model.processer_list, state, x0, e0, bufs, row_enc, col_enc, model.prec)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x.cpu().numpy(), e=e.cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("name", ["MgnA", "FvgnA"])
def test_peer_memory_gather_equals_single_gpu(tmp_path, name):
    """Ghost rows read straight from the owner's HBM over NVLink inside the fused edge kernel: bit-identical."""
    from gnn_fluid_dynamics_b200.topology import get_topology
    world = min(torch.cuda.device_count(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_peer_rank, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    model, graphs, parts, _ = _setup(name, 4000, world, dev)
    gd = [g.to(dev) for g in graphs]
    from gnn_fluid_dynamics_b200 import processor as P
    with torch.no_grad(), P.no_fast():     # the peer-memory gathers exist in the register-staged kernels only
        x, e, _ = model.encode_process_decode(gd[0].x, gd[1].x, get_topology(gd).validate())
    for p in parts:
        d = np.load(tmp_path / f"rank{p.rank}.npz")
        assert np.array_equal(d["x"], x.cpu()[p.cells[:p.n_owned]].numpy())
        assert np.array_equal(d["e"], e.cpu()[p.faces].numpy())


def test_partition_invariance_and_determinism_at_200k_cells():
    """BASELINE.json config 3 size (200k cells): size-independent properties instead of an oracle run - the result
    does not depend on how the mesh is cut (4 partitions == unpartitioned, bit for bit) nor on the run."""
    from gnn_fluid_dynamics_b200.dist import InProcessTransport, encode_process_decode_partitioned
    from gnn_fluid_dynamics_b200.topology import get_topology
    dev = torch.device("cuda:0")
    model, graphs, parts, locals_ = _setup("MgnA", 200000, 4, dev)
    gd = [g.to(dev) for g in graphs]
    with torch.no_grad():
        topo = get_topology(gd).validate()
        x, e, dec = model.encode_process_decode(gd[0].x, gd[1].x, topo)
        x2, e2, dec2 = model.encode_process_decode(gd[0].x, gd[1].x, topo)
        assert torch.equal(x, x2) and torch.equal(e, e2) and torch.equal(dec, dec2)
        assert torch.isfinite(x).all() and torch.isfinite(e).all()
        states, inputs = _states(parts, locals_, dev)
        outs = encode_process_decode_partitioned(model, states, inputs, InProcessTransport())
    for p, (xp, ep, dp) in zip(parts, outs):
        assert torch.equal(xp, x[p.cells[:p.n_owned].to(dev)])
        assert torch.equal(ep, e[p.faces.to(dev)])
        assert torch.equal(dp, dec[p.cells[:p.n_owned].to(dev)])
