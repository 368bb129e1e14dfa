"""Host-side (Python) cost of enqueuing one training step: cProfile over 5 steps, top functions by own time."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_model
from bench import build_samples, _collate
from gnn_fluid_dynamics_b200.topology import get_topology
dev = torch.device("cuda:0")
model = build_model("FvgnA", precision="bf16x3").to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
gd = [g.to(dev) for g in _collate(build_samples("FvgnA", 8, 20000, "cylinder"))]
gd = model.normalizer.input(gd)
topo = get_topology(gd).validate(); topo.build_row_col_interleaved_csr(); topo.build_vf_csr()
gd[0].topology = gd[2].topology = topo
def step():
    opt.zero_grad(set_to_none=True)
    out = model.forward_normalised(gd, mode="train")
    loss = model.loss(out, gd)["total_log_loss"]
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    step(); torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
