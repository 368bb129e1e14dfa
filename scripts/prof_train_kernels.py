"""One forward (+stash) and one backward of the fused edge MLP at the bench shape, twice (first = warm-up), so
`ncu -k regex:"wgrad_tc|mlp_tc_kernel" -s 7 -c 7` captures exactly: forward+stash, wgrad L3, dgrad chain,
wgrad L2, wgrad L1 (gathered concat), 2 x single-Linear dIn."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnn_fluid_dynamics_b200 import ops, _lib, training
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
x = torch.randn(N, 128, generator=g).to(dev); e = torch.randn(E, 128, generator=g).to(dev); go = torch.randn(E, 128, generator=g).to(dev)
we = _to_weights(_rand_mlp(384, 128, True, seed=1), 0)
segs = [Seg(e), Seg(x, _lib.SEG_GATHER, (row,)), Seg(x, _lib.SEG_GATHER, (col,))]
ws = ops.mlp_backward_workspace(E, dev)
flush = torch.empty(200 * 1024 * 1024 // 4, device=dev)
for it in range(2):
    flush.zero_()
    _, _, st = ops.mlp_forward(segs, we, E, _lib.PREC_BF16X3, residual=e, want_raw=False, want_sum=True, stash=True)
    training.mlp_backward(we, st, segs, E, go, _lib.PREC_BF16X3, [{"residual": go}, {}, {}], ws)
torch.cuda.synchronize()
print("E", E, "N", N)
