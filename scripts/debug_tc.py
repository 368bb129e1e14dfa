"""Ad-hoc numerics probe for the tensor-core MLP variants (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import oracle
from gnn_fluid_dynamics_b200 import ops, _lib
from gnn_fluid_dynamics_b200.testing import rel_l2
from test_gpu_parity import _rand_mlp, _to_weights

dev = torch.device("cuda:0")
precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16x1", "bf16x3", "fp16x2", "fp16x3"]
cases = [(128, 128, True, 128), (128, 128, True, 1000), (384, 128, True, 777), (192, 128, True, 300),
         (10, 128, True, 200), (2, 128, True, 129), (128, 5, False, 500), (256, 128, True, 4096)]
for prec in precs:
    for k_in, n_out, ln, rows in cases:
        p = _rand_mlp(k_in, n_out, ln, True, seed=k_in + rows)
        x = torch.randn(rows, k_in, generator=torch.Generator().manual_seed(1))
        ref = oracle.mlp3(x, p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["ln_w"], p["ln_b"])
        w = _to_weights(p, 0)
        out, _ = ops.mlp_forward([ops.Seg(x.to(dev))], w, rows, _lib.PRECISIONS[prec])
        torch.cuda.synchronize()
        print(f"{prec:8s} k_in={k_in:4d} n_out={n_out:4d} rows={rows:5d} rel_l2={rel_l2(out, ref):.3e} "
              f"finite={bool(torch.isfinite(out).all())}", flush=True)
