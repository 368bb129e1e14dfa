"""GraphCache: device-side collation == the host collation of the reference's loader
(graph.collate_triplet restates PyG Batch.from_data_list), static attributes uploaded once, dynamic ones per fetch."""
import pytest
import torch

from gnn_fluid_dynamics_b200.graph import collate_triplet
from gnn_fluid_dynamics_b200.graph_cache import GraphCache
from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs


def _samples(seed0=0, n=3, flip=True):
    return [mesh_graphs(make_mesh(300 + 150 * i, "cylinder", seed=i), seed=seed0 + i, flip_edges=flip) for i in range(n)]


def _same(ref, out):
    for gi in range(3):
        keys = set(ref[gi].keys()) | set(out[gi].keys())
        for k in sorted(keys):
            a, b = ref[gi]._store.get(k), out[gi]._store.get(k)
            assert a is not None and b is not None, (gi, k)
            if torch.is_tensor(a):
                assert a.dtype == b.dtype and torch.equal(a, b.cpu()), (gi, k)
        assert ref[gi].num_nodes == out[gi].num_nodes


def test_device_collation_equals_host_collation_and_refreshes_dynamic_attributes():
    cache = GraphCache("cpu")
    s0 = _samples(seed0=10)
    _same(collate_triplet(s0), cache.fetch([0, 1, 2], s0))
    first = cache.h2d_bytes
    assert cache.mesh_misses == 3 and cache.batch_misses == 1
    # another timestep of the same meshes: new state / targets / orientation, same geometry
    s1 = _samples(seed0=50)
    out = cache.fetch([0, 1, 2], s1)
    _same(collate_triplet(s1), out)
    assert cache.batch_hits == 1 and cache.h2d_static_bytes == 0 and 0 < cache.h2d_bytes < first
    # another batch composition of meshes already resident: nothing static travels
    out = cache.fetch([2, 0], [s1[2], s1[0]])
    _same(collate_triplet([s1[2], s1[0]]), out)
    assert cache.mesh_hits >= 2 and cache.h2d_static_bytes == 0


def test_shape_change_under_the_same_key_is_refused():
    cache = GraphCache("cpu")
    s0 = _samples()
    cache.fetch([0, 1, 2], s0)
    bad = [s0[1], s0[1], s0[2]]
    with pytest.raises(RuntimeError):
        cache.fetch([0, 1, 2], bad)


@pytest.mark.gpu
def test_training_step_through_the_cache_equals_the_direct_path():
    from helpers import build_model, LOSS_W  # noqa: F401
    dev = torch.device("cuda:0")
    model = build_model("FvgnA", mp_num=3).to(dev).train()
    samples = [mesh_graphs(make_mesh(500 + 60 * i, "cylinder", seed=i), seed=20 + i, flip_edges=True) for i in range(3)]
    for g in samples:
        g[1].y = g[1].y[:, :3].contiguous()

    def grads(graphs):
        model.zero_grad(set_to_none=True)
        gn = model.normalizer.input(graphs)
        out = model.forward_normalised(gn, mode="train")
        model.loss(out, gn)["total_log_loss"].backward()
        return [p.grad.clone() for p in model.parameters() if p.grad is not None]

    ref = grads([g.to(dev) for g in collate_triplet([[t.clone() for t in s] for s in samples])])
    cache = GraphCache(dev)
    for it in range(2):          # second fetch: topology attached by the first step, orientation refreshed in place
        got = grads(cache.fetch([0, 1, 2], samples))
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), it
    for it in range(3):          # double-buffered: the next batch is uploaded on the copy stream while this one trains
        g = cache.fetch([0, 1, 2], samples)
        cache.prefetch([0, 1, 2], samples)
        got = grads(g)
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), ("prefetch", it)


@pytest.mark.gpu
def test_refresh_orientation_equals_fresh_build():
    from gnn_fluid_dynamics_b200.topology import MeshTopology
    dev = torch.device("cuda:0")
    g = [t.to(dev) for t in mesh_graphs(make_mesh(2000, "cylinder", seed=3), seed=4, flip_edges=False)]
    topo = MeshTopology.from_graphs(g).validate()
    topo.build_rowcol_csr(); topo.build_row_csr(); topo.build_cell_csr(); topo.build_row_col_interleaved_csr()
    ei = g[0].edge_index.clone()
    flip = torch.rand(ei.shape[1], device=dev) < 0.5
    ei[:, flip] = ei[:, flip].flip(0)
    g[0].edge_index = ei
    topo.refresh_orientation(ei).validate()
    fresh = MeshTopology.from_graphs(g).validate()
    assert torch.equal(topo.row, fresh.row) and torch.equal(topo.col, fresh.col)
    for name in ("build_rowcol_csr", "build_row_csr", "build_col_csr", "build_cell_csr", "build_row_col_interleaved_csr"):
        a, b = getattr(topo, name)(), getattr(fresh, name)()
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), name
    assert torch.equal(topo.vtx_perm, fresh.vtx_perm)
