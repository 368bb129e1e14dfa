"""Inference node block (fast path) at the bench shape, 3 launches (ncu: -k regex:mlp_tc_kernel -s 2 -c 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnn_fluid_dynamics_b200 import ops, _lib
from gnn_fluid_dynamics_b200.ops import Seg
from gnn_fluid_dynamics_b200.mesh import make_mesh
from test_gpu_parity import _rand_mlp, _to_weights

dev = torch.device("cuda:0")
meshes = [make_mesh(20000, "cylinder", seed=i) for i in range(8)]
vfs, off, voff = [], 0, 0
for m in meshes:
    vfs.append(torch.from_numpy(m.cells).T.contiguous() + voff)
    off += m.n_cells; voff += m.n_vertices
vf = tuple(t.to(torch.int32).to(dev).contiguous() for t in torch.cat(vfs, 1))
N, V = off, voff
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 128, generator=g).to(dev); vs = torch.randn(V, 64, generator=g).to(dev)
wn = _to_weights(_rand_mlp(192, 128, True, seed=2), 0)
xs = torch.empty(N, 256, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
nsegs = [Seg(x), Seg(vs, _lib.SEG_MEAN3, vf)]
for it in range(3):
    flush.zero_()
    ops.mlp_forward(nsegs, wn, N, _lib.PREC_BF16X3, residual=x, want_raw=False, want_sum=True, out_sum=x, out_split=xs)
torch.cuda.synchronize()
print("N", N, "V", V)
