"""Scratch experiment: edge-block forward time vs. gather locality (real mesh indices / small range / identity)."""
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
ar = torch.arange(E, device=dev, dtype=torch.int32) % N
cases = {
    "mesh": (row, col), "small": (row % 4096, col % 4096), "identity": (ar, ar),
    "sorted": (torch.sort(row).values.contiguous(), torch.sort(col).values.contiguous()),
}
def run(segs, prec):
    ts = []
    for it in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.mlp_forward(segs, we, E, prec, residual=e, want_raw=False, want_sum=True)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts[2:])
for prec in ("bf16x3", "bf16x1"):
    P = _lib.PRECISIONS[prec]
    for name, (r, c) in cases.items():
        segs = [Seg(e), Seg(x, _lib.SEG_GATHER, (r,)), Seg(x, _lib.SEG_GATHER, (c,))]
        print(f"{prec} {name:9s} {run(segs, P):7.1f} us")

