"""Timing ablations of the fused edge / node blocks at the bench shape (8 x 20k-cell meshes).  Run once per library
build:  GNNFD_LIB=gnn_fluid_dynamics_b200/lib_abl/libgnnfd_abl3.so python scripts/abl_edge.py <tag>
(builds with -DGNNFD_ABL=n compute WRONG results; they only say what each part of the kernel costs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnn_fluid_dynamics_b200 import ops, _lib
from gnn_fluid_dynamics_b200.ops import Seg
from gnn_fluid_dynamics_b200.mesh import make_mesh
from test_gpu_parity import _rand_mlp, _to_weights

tag = sys.argv[1] if len(sys.argv) > 1 else "base"
dev = torch.device("cuda:0")
meshes = [make_mesh(20000, "cylinder", seed=i) for i in range(8)]
rows, cols, vfs, off, voff = [], [], [], 0, 0
for m in meshes:
    rows.append(torch.from_numpy(m.cell_edge_index[0]) + off); cols.append(torch.from_numpy(m.cell_edge_index[1]) + off)
    vfs.append(torch.from_numpy(m.cells).T.contiguous() + voff)
    off += m.n_cells; voff += m.n_vertices
row, col = torch.cat(rows).to(torch.int32).to(dev), torch.cat(cols).to(torch.int32).to(dev)
vf = tuple(t.to(torch.int32).to(dev).contiguous() for t in torch.cat(vfs, 1))
N, E, V = off, row.numel(), voff
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 128, generator=g).to(dev); e = torch.randn(E, 128, generator=g).to(dev)
vs = torch.randn(V, 64, generator=g).to(dev)
we = _to_weights(_rand_mlp(384, 128, True, seed=1), 0); wn = _to_weights(_rand_mlp(192, 128, True, seed=2), 0)
esegs = [Seg(e), Seg(x, _lib.SEG_GATHER, (row,)), Seg(x, _lib.SEG_GATHER, (col,))]
nsegs = [Seg(x), Seg(vs, _lib.SEG_MEAN3, vf)]
P = _lib.PREC_BF16X3
big = [torch.randn(E, 128, device=dev) for _ in range(3)]   # > L2 of other traffic between launches


def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for i in range(n):
        big[i % 3].add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


te = timeit(lambda: ops.mlp_forward(esegs, we, E, P, residual=e, want_raw=False, want_sum=True))
tn = timeit(lambda: ops.mlp_forward(nsegs, wn, N, P, residual=x, want_raw=True, want_sum=True))
print(f"{tag:8s} edge {te[0]:7.1f} us (min {te[1]:7.1f})   node {tn[0]:7.1f} us (min {tn[1]:7.1f})   E={E} N={N} V={V}", flush=True)
# inference fast path: TMA-gathered split shadow, in-place residual (TMA reduce-add), shadow written by the node block
xs = torch.empty(N, 256, dtype=torch.bfloat16, device=dev)
hi = x.to(torch.bfloat16); xs[:, :128] = hi; xs[:, 128:] = (x - hi.float()).to(torch.bfloat16)
xf = xs.view(torch.float32)
fsegs = [Seg(e), Seg(xf, _lib.SEG_GATHER, (row,), split=xs), Seg(xf, _lib.SEG_GATHER, (col,), split=xs)]
tef = timeit(lambda: ops.mlp_forward(fsegs, we, E, P, residual=e, want_raw=False, want_sum=True, out_sum=e))
# in-place epilogue only (register-staged gathers)
tei = timeit(lambda: ops.mlp_forward(esegs, we, E, P, residual=e, want_raw=False, want_sum=True, out_sum=e))
tnf = timeit(lambda: ops.mlp_forward(nsegs, wn, N, P, residual=x, want_raw=False, want_sum=True, out_sum=x, out_split=xs))
alg_e, alg_n = (512 * (2 * E + N) + 8 * E) / 1e3, (512 * 3 * N + 256 * V + 12 * N) / 1e3
print(f"{tag:8s} FAST edge {tef[0]:7.1f} us (min {tef[1]:7.1f}, {alg_e / tef[0]:6.0f} GB/s)   edge in-place epilogue only {tei[0]:7.1f} us"
      f"   node {tnf[0]:7.1f} us (min {tnf[1]:7.1f}, {alg_n / tnf[0]:6.0f} GB/s)", flush=True)
