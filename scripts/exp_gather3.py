import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnn_fluid_dynamics_b200 import fvm_ops
from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs
from gnn_fluid_dynamics_b200.graph import collate_triplet
from gnn_fluid_dynamics_b200.topology import get_topology
dev = torch.device("cuda:0")
g = collate_triplet([mesh_graphs(make_mesh(20000, "cylinder", seed=i), seed=i) for i in range(8)])
gd = [x.to(dev) for x in g]
topo = get_topology(gd).validate()
N, E = gd[0].x.shape[0], gd[0].edge_index.shape[1]
cf = gd[1].face
cfs = fvm_ops.cell_faces(topo, cf)
def timeit(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:40s} gpu {a.elapsed_time(b)/n*1e3:9.1f} us   wall {(time.perf_counter()-t0)/n*1e6:9.1f} us", flush=True)
for w in (2, 5):
    t = torch.randn(E, w, device=dev, requires_grad=True)
    gg = torch.randn(3, N, w, device=dev)
    timeit(f"gather3 fwd w={w}", lambda: fvm_ops.gather3(t.detach(), cfs, topo.row, topo.col))
    def fb():
        t.grad = None
        (fvm_ops.gather3(t, cfs, topo.row, topo.col) * gg).sum().backward()
    timeit(f"gather3 fwd+bwd w={w}", fb)
    def fb2():
        t.grad = None
        (torch.stack([t[cf[0]], t[cf[1]], t[cf[2]]]) * gg).sum().backward()
    timeit(f"indexing fwd+bwd w={w}", fb2)
