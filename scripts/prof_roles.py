"""Role wait-time breakdown of the tensor-core MLP kernel (CTA 0), edge/node block shapes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import build_model
from gnn_fluid_dynamics_b200 import processor as P, _lib
from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs
from gnn_fluid_dynamics_b200.graph import collate_triplet
from gnn_fluid_dynamics_b200.topology import get_topology

dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
model = build_model("FvgnA", precision=prec).to(dev).eval()
# optional: cells per mesh, number of meshes (default = the bench shape; "2000 1" = the launch-bound 2k-cell rollout shape)
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n_meshes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = collate_triplet([mesh_graphs(make_mesh(n_cells, "cylinder", seed=i), seed=i) for i in range(n_meshes)])
gd = [x.to(dev) for x in g]
topo = get_topology(gd).validate()
N, E = gd[0].x.shape[0], gd[0].edge_index.shape[1]
x = torch.randn(N, 128, device=dev); e = torch.randn(E, 128, device=dev)
blk = model.processer_list[3]
vs = P.vertex_half_sum(e, topo)
fast = P.Fast(N, model.prec, dev)
hi = x.to(fast.dtype); fast.xs[:, :128] = hi; fast.xs[:, 128:] = (x - hi.float()).to(fast.dtype)
torch.set_grad_enabled(False)
for which in ("edge", "node", "edge_fast", "node_fast"):
    def run():
        if which == "edge":
            P.edge_mlp_concat(blk.face_block.face_mlp, e, x, topo, model.prec, want_raw=False)
        elif which == "node":
            P.node_mlp_two_hop(blk.cell_block.cell_mlp, x, vs, topo, model.prec, want_raw=True)
        elif which == "edge_fast":
            P.edge_mlp_concat(blk.face_block.face_mlp, e, None, topo, model.prec, want_raw=False, fast=fast)
        else:
            P.node_mlp_two_hop(blk.cell_block.cell_mlp, x, vs, topo, model.prec, want_raw=False, fast=fast)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 16)()
    _lib.check(_lib.lib.gnnfd_tc_profile_read(buf), "profile")
    v = list(buf)
    T = max(v[6], 1)
    print(f"[{which} {prec}] kernel {a.elapsed_time(b)*1e3:.1f} us, tiles/CTA {T}")
    print(f"  L1 issuer  : total {v[0]/T:8.0f} cyc/tile | w_full {v[2]/T:7.0f} acc_free {v[3]/T:7.0f} a_full {v[4]/T:7.0f}")
    print(f"  L23 issuer : w23_full {v[1]/T:7.0f} hid_ready {v[5]/T:7.0f}")
    print(f"  epilogue g0: (its own tiles = every other one; per-tile figures below are over ALL tiles)")
    print(f"  epilogue   : total {v[8]/T:8.0f} cyc/tile | wait hidden {v[9]/T:7.0f} wait final {v[10]/T:7.0f}")
    print(f"  producer   : total {v[12]/T:8.0f} cyc/tile | wait a_empty {v[13]/T:7.0f} convert+store {v[14]/T:7.0f} fence+arrive {v[15]/T:7.0f} issue {v[7]/T:7.0f} boundary {v[11]/T:7.0f}")
