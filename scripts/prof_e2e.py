"""Where the end-to-end training step's time goes on the host: per-phase wall clock of bench.py's e2e loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from helpers import build_model
from gnn_fluid_dynamics_b200.graph_cache import GraphCache

dev = torch.device("cuda:0")
model = build_model("FvgnA", precision="bf16x3").to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
samples = [[g.pin_memory() for g in s] for s in bench.build_samples("FvgnA", 8, 20000, "cylinder")]
cache = GraphCache(dev)
keys = list(range(8))
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(measure):
    t = time.perf_counter(); g = cache.fetch(keys, samples); measure and tick("fetch", t)
    before = torch.cuda.Event(); before.record()
    t = time.perf_counter(); gn = model.normalizer.input(g); measure and tick("normalizer.input", t)
    t = time.perf_counter(); opt.zero_grad(set_to_none=True); out = model.forward_normalised(gn, mode="train"); measure and tick("forward (enqueue)", t)
    t = time.perf_counter(); loss = model.loss(out, gn)["total_log_loss"]; measure and tick("loss (enqueue)", t)
    t = time.perf_counter(); loss.backward(); measure and tick("backward (enqueue)", t)
    t = time.perf_counter(); torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); opt.step(); measure and tick("clip + Adam (enqueue)", t)
    t = time.perf_counter(); cache.prefetch(keys, samples, after=before); measure and tick("prefetch (enqueue)", t)
    t = time.perf_counter(); v = float(loss.item()); measure and tick("loss.item() (wait for the GPU)", t)
for _ in range(4): step(False)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n): step(True)
tot = (time.perf_counter() - t0) / n * 1e3
print(f"e2e step {tot:.2f} ms")
for k, v in T.items(): print(f"  {k:34s} {v / n * 1e3:7.3f} ms")
