"""Domain decomposition plan (host logic) + the world_size-2 halo exchange over gloo, on CPU.

The compute of each rank is the CPU oracle on the rank's LOCAL sub-mesh (tests may use the oracle); what is under
test is the product's partition plan (gnn_fluid_dynamics_b200/partition.py) and transport
(gnn_fluid_dynamics_b200/dist.py: TorchDistTransport), which must reproduce the unpartitioned result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import oracle
from oracle.blocks import face_block_concat, two_hop_aggregate
from oracle.mlp import mlp_from_state
from gnn_fluid_dynamics_b200.partition import local_graphs, partition_mesh
from helpers import build_model, golden_graphs


def _mesh(n_cells=600, name="MgnA"):
    mesh, graphs = golden_graphs(name, n_cells=n_cells, mesh_seed=11, feat_seed=12)
    return mesh, graphs


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_plan_invariants(world):
    mesh, graphs = _mesh(900)
    c, f, v = graphs
    parts = partition_mesh(c.edge_index, v.edge_index, v.face, c.pos[:, 0], world, f_face=f.face)
    N, E = c.x.shape[0], c.edge_index.shape[1]
    owned_all = torch.cat([p.cells[:p.n_owned] for p in parts])
    assert torch.equal(torch.sort(owned_all).values, torch.arange(N))          # owners partition the cells
    assert max(p.n_owned for p in parts) - min(p.n_owned for p in parts) <= 1      # equal-count strips
    vf, vei, cei = v.face, v.edge_index, c.edge_index
    for p in parts:
        owned = p.cells[:p.n_owned]
        assert torch.equal(owned, torch.sort(owned).values)
        # ghost cells == vertex-star of the owned cells minus the owned cells (brute force)
        own_v = torch.unique(vf[:, owned])
        star = torch.nonzero(torch.isin(vf, own_v).any(0)).flatten()
        ghosts = torch.sort(p.cells[p.n_owned:]).values
        expect = star[~torch.isin(star, owned)]
        assert torch.equal(ghosts, expect)
        # local faces == faces touching an owned vertex; every local face has both cells local
        fexp = torch.nonzero(torch.isin(vei, own_v).any(0)).flatten()
        assert torch.equal(p.faces, fexp)
        assert torch.equal(p.cells[p.c_edge_index], cei[:, p.faces])
        assert torch.equal(p.verts[p.v_edge_index], vei[:, p.faces])
        assert torch.equal(p.verts[p.v_face], vf[:, owned])
        assert torch.equal(p.faces[p.f_face], f.face[:, owned])
        # receive ranges tile the ghost rows in owner order; send lists mirror the peers' receive lists
        pos = p.n_owned
        for peer in sorted(p.recv):
            start, cnt = p.recv[peer]
            assert start == pos and cnt > 0 and peer != p.rank
            pos += cnt
            want = p.cells[start:start + cnt]
            q = parts[peer]
            assert torch.equal(q.cells[q.send[p.rank]], want)
        assert pos == p.n_local


def _oracle_block_local(family, sd, i, x, e, g, n_owned, exchange):
    """One GN_Block on a local sub-mesh with the oracle's sub-blocks; node phase on owned cells only."""
    c, f, v = g
    p = f"processer_list.{i}"
    if family == "mgn":
        if i > 0:
            exchange(x)
        er = face_block_concat(sd, f"{p}.face_block.face_mlp", x, e, c.edge_index)
        agg, _ = two_hop_aggregate(er, v.edge_index, v.face, v.pos.shape[0])
        xr = mlp_from_state(sd, f"{p}.cell_block.cell_mlp", torch.cat([x[:n_owned], agg], -1))
        x = x.clone()
        x[:n_owned] += xr
        return x, e + er
    agg, _ = two_hop_aggregate(e, v.edge_index, v.face, v.pos.shape[0])
    xr = torch.zeros_like(x)
    xr[:n_owned] = mlp_from_state(sd, f"{p}.cell_block.cell_mlp", torch.cat([x[:n_owned], agg], -1))
    exchange(xr)
    er = face_block_concat(sd, f"{p}.face_block.face_mlp", xr, e, c.edge_index)
    x = x.clone()
    x[:n_owned] += xr[:n_owned]
    return x, e + er


def _rank_main(rank, world, port, name, family, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from gnn_fluid_dynamics_b200.dist import PartState, TorchDistTransport
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        model = build_model(name, mp_num=4)
        sd = model.state_dict()
        _, graphs = _mesh(600, name)
        graphs = model.normalizer.input([g.clone() for g in graphs])
        c, f, v = graphs
        parts = partition_mesh(c.edge_index, v.edge_index, v.face, c.pos[:, 0], world)
        part = parts[rank]
        g = local_graphs(graphs, part)
        state = PartState(part=part, topo=None)
        transport = TorchDistTransport(pack=lambda t, idx: t.index_select(0, idx.long()))
        with torch.no_grad():
            x, e, _ = oracle.encoder_fwd(family, sd, g[0].x, g[1].x)
            for i in range(4):
                x, e = _oracle_block_local(family, sd, i, x, e, g, part.n_owned,
                                           lambda t: transport.exchange([state], lambda s: t))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x[:part.n_owned].numpy(), e=e.numpy(),
                 cells=part.cells[:part.n_owned].numpy(), faces=part.faces.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("name,family", [("MgnA", "mgn"), ("FvgnA", "fvgn")])
def test_world2_gloo_halo_exchange_matches_unpartitioned(tmp_path, name, family):
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), name, family, str(tmp_path)), nprocs=world, join=True)
    model = build_model(name, mp_num=4)
    sd = model.state_dict()
    _, graphs = _mesh(600, name)
    graphs = model.normalizer.input([g.clone() for g in graphs])
    c, f, v = graphs
    topo = {"c_edge_index": c.edge_index, "v_edge_index": v.edge_index, "v_face": v.face, "n_vertices": v.num_nodes}
    with torch.no_grad():
        ref = oracle.processor_fwd(family, sd, c.x, f.x, topo, 4)
    seen = torch.zeros(c.x.shape[0], dtype=torch.bool)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        cells, faces = torch.from_numpy(d["cells"]), torch.from_numpy(d["faces"])
        assert torch.allclose(torch.from_numpy(d["x"]), ref["x"][cells], rtol=0, atol=2e-6)
        assert torch.allclose(torch.from_numpy(d["e"]), ref["e"][faces], rtol=0, atol=2e-6)
        seen[cells] = True
    assert bool(seen.all())


def _dp_rank(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from gnn_fluid_dynamics_b200.dist import allreduce_gradients
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
        g = torch.Generator().manual_seed(100 + rank)
        params[0].grad = torch.randn(5, 3, generator=g)
        params[1].grad = torch.randn(7, generator=g)      # params[2] has no gradient (unused parameter)
        allreduce_gradients(params, world)
        np.savez(os.path.join(out_dir, f"dp{rank}.npz"), g0=params[0].grad.numpy(), g1=params[1].grad.numpy())
    finally:
        dist.destroy_process_group()


def test_world2_gloo_data_parallel_gradient_allreduce(tmp_path):
    """Independent meshes per rank (BASELINE.json configs 2 / 5): one flat all-reduce averages every gradient."""
    world = 2
    mp.spawn(_dp_rank, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    exp0 = sum(torch.randn(5, 3, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for g in gens:
        torch.randn(5, 3, generator=g)
    exp1 = sum(torch.randn(7, generator=g) for g in gens) / world
    for r in range(world):
        d = np.load(tmp_path / f"dp{r}.npz")
        assert np.allclose(d["g0"], exp0.numpy(), atol=1e-7) and np.allclose(d["g1"], exp1.numpy(), atol=1e-7)
