"""Inference forward of the fused edge block at the bench shape, 3 launches (ncu: -k regex:mlp_tc_kernel -s 2 -c 1)."""
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
rows, cols, off = [], [], 0
for m in meshes:
    rows.append(torch.from_numpy(m.cell_edge_index[0]) + off); cols.append(torch.from_numpy(m.cell_edge_index[1]) + off)
    off += m.n_cells
row, col = torch.cat(rows).to(torch.int32).to(dev), torch.cat(cols).to(torch.int32).to(dev)
N, E = off, row.numel()
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 128, generator=g).to(dev); e = torch.randn(E, 128, generator=g).to(dev)
we = _to_weights(_rand_mlp(384, 128, True, seed=1), 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
fast = len(sys.argv) > 1 and sys.argv[1] == "fast"      # inference fast path: TMA-gathered split shadow, in-place residual
if fast:
    xs = torch.empty(N, 256, dtype=torch.bfloat16, device=dev)
    hi = x.to(torch.bfloat16); xs[:, :128] = hi; xs[:, 128:] = (x - hi.float()).to(torch.bfloat16)
    xf = xs.view(torch.float32)
    segs = [Seg(e), Seg(xf, _lib.SEG_GATHER, (row,), split=xs), Seg(xf, _lib.SEG_GATHER, (col,), split=xs)]
else:
    segs = [Seg(e), Seg(x, _lib.SEG_GATHER, (row,)), Seg(x, _lib.SEG_GATHER, (col,))]
for it in range(3):
    flush.zero_()
    ops.mlp_forward(segs, we, E, _lib.PREC_BF16X3, residual=e, want_raw=False, want_sum=True, out_sum=e if fast else None)
torch.cuda.synchronize()
print("E", E, "N", N)
