"""Per-kernel timings at the bench shape (E=242k, N=160k, V=82k) with CUDA events; L2 flushed between launches.
   python scripts/bench_kernels.py  -> one line per kernel: ms, algorithmic MB, GB/s"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gnn_fluid_dynamics_b200 import ops, _lib, training
from gnn_fluid_dynamics_b200.ops import Seg
from test_gpu_parity import _rand_mlp, _to_weights

dev = torch.device("cuda:0")
E, N, V = 242328, 160327, 82001
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 128, generator=g).to(dev); e = torch.randn(E, 128, generator=g).to(dev)
vs = torch.randn(V, 64, generator=g).to(dev)
i32 = lambda t: t.to(torch.int32).to(dev)
# mesh-like locality: neighbours are near in index space
base = torch.arange(E) * N // E
row = i32((base + torch.randint(-200, 200, (E,), generator=g)).clamp(0, N - 1))
col = i32((base + torch.randint(-200, 200, (E,), generator=g)).clamp(0, N - 1))
vb = torch.arange(N) * V // N
vf = tuple(i32((vb + torch.randint(-150, 150, (N,), generator=g)).clamp(0, V - 1)) for _ in range(3))
flush = torch.empty(200 * 1024 * 1024 // 4, device=dev)
P = _lib.PREC_BF16X3

def timeit(name, fn, mbytes, n=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {ms*1e3:8.1f} us  {mbytes:8.1f} MB  {mbytes/ms:8.1f} GB/s", flush=True)

we = _to_weights(_rand_mlp(384, 128, True, seed=1), 0); wn = _to_weights(_rand_mlp(192, 128, True, seed=2), 0)
esegs = [Seg(e), Seg(x, _lib.SEG_GATHER, (row,)), Seg(x, _lib.SEG_GATHER, (col,))]
nsegs = [Seg(x), Seg(vs, _lib.SEG_MEAN3, vf)]
MB = 1e-6
timeit("fwd edge MLP (inference)", lambda: ops.mlp_forward(esegs, we, E, P, residual=e, want_raw=False, want_sum=True), (512*(2*E+N)+8*E)*MB)
timeit("fwd edge MLP (+stash a1,a2,xhat)", lambda: ops.mlp_forward(esegs, we, E, P, residual=e, want_raw=False, want_sum=True, stash=True), (512*(5*E+N)+8*E)*MB)
timeit("fwd node MLP (inference, raw+sum)", lambda: ops.mlp_forward(nsegs, wn, N, P, residual=x, want_raw=True, want_sum=True), (512*3*N+256*V+12*N)*MB)
timeit("fwd node MLP (+stash)", lambda: ops.mlp_forward(nsegs, wn, N, P, residual=x, want_raw=True, want_sum=True, stash=True), (512*6*N+256*V+12*N)*MB)
_, _, ste = ops.mlp_forward(esegs, we, E, P, residual=e, want_raw=False, want_sum=True, stash=True)
_, _, stn = ops.mlp_forward(nsegs, wn, N, P, residual=x, want_raw=True, want_sum=True, stash=True)
go = torch.randn(E, 128, generator=g).to(dev); gn = go[:N].contiguous()
ws = ops.mlp_backward_workspace(E, dev)
timeit("edge MLP backward (fused call, 3 dIn)", lambda: training.mlp_backward(we, ste, esegs, E, go, P, [{"residual": go}, {}, {}], ws), 0.0)
timeit("node MLP backward (fused call, 2 dIn)", lambda: training.mlp_backward(wn, stn, nsegs, N, gn, P, [{"residual": gn}, {}], ws), 0.0)
timeit("ln_backward E", lambda: ops.ln_backward(go, ste.xhat, ste.rstd, we.ln_w), 512*3*E*MB)
packs = {}
timeit("single linear E (mul act')", lambda: ops.linear_tc(Seg(go), E, we.w2, 1, 128, 128, 128, packs, "a", P, mul=ste.a1, mul_mode=1), 512*3*E*MB)
timeit("single linear E (plain)", lambda: ops.linear_tc(Seg(go), E, we.w2, 1, 128, 128, 128, packs, "a", P), 512*2*E*MB)
out = torch.empty(128, 128, device=dev); out3 = torch.empty(128, 384, device=dev); cs = torch.empty(128, device=dev)
timeit("wgrad E n=128 (B=silu(a))", lambda: ops.wgrad(Seg(go), [Seg(ste.a1)], E, out, b_act=1, colsum=cs), 512*2*E*MB)
timeit("wgrad E n=128 (no act)", lambda: ops.wgrad(Seg(go), [Seg(ste.a1)], E, out), 512*2*E*MB)
timeit("wgrad E n=384 (gathered concat)", lambda: ops.wgrad(Seg(go), esegs, E, out3, colsum=cs), (512*(2*E+N)+8*E)*MB)
# chain only
import ctypes as C
def chain():
    b = _lib.MlpBackwardArgs()
    return None
rc = ops.csr_build(torch.cat([row, col]), N)
t1 = torch.randn(E, 128, generator=g).to(dev); t2 = torch.randn(E, 128, generator=g).to(dev)
timeit("segment_sum3 rowcol (E->N, w128)", lambda: ops.segment_sum3(t1, t2, None, (0, 0, 0), 128, 1.0, E, rc[0], rc[1], N, base=x), (512*(2*E+2*N)+8*E)*MB)
v0 = i32(torch.randint(0, V, (E,), generator=g)); v1 = i32(torch.randint(0, V, (E,), generator=g))
timeit("gather_pair_add halves (V->E)", lambda: ops.gather_pair_add(vs, v0, v1, 1.0, True, E, base=t1, out=t1), (512*2*E+256*V)*MB)
timeit("segment_sum halves (E->V)", lambda: ops.segment_sum(e, e, 0, 64, 64, 1.0, *ops.csr_build(torch.cat([v0, v1]), V), V), (512*E+256*V)*MB, n=4)
