import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from gnn_fluid_dynamics_b200 import ops, _lib
from gnn_fluid_dynamics_b200.dist import PeerBuffers, PEER_SHIFT
from gnn_fluid_dynamics_b200.ops import Seg
bufs = PeerBuffers(1000, 128, dev, world, rank)
bufs.local[0].fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
v = bufs.views[0][peer]
print(rank, "peer view device", v.device, "ptr", hex(v.data_ptr()), "can access", torch.cuda.can_device_access_peer(rank, peer), flush=True)
print(rank, "torch read of peer:", float(v[:4, :4].sum()), flush=True)
idx = torch.arange(0, 1000, 7, dtype=torch.int32, device=dev)
out = ops.gather_rows(v, idx); torch.cuda.synchronize()
print(rank, "gather_rows over P2P:", float(out.mean()), flush=True)
# MLP kernel with peer gather
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import _rand_mlp
from gnn_fluid_dynamics_b200.ops import MLPWeights
p = _rand_mlp(256, 128, True, seed=1)
d = lambda t: None if t is None else t.to(dev).contiguous()
w = MLPWeights(w1=d(p["w1"]), b1=d(p["b1"]), w2=d(p["w2"]), b2=d(p["b2"]), w3=d(p["w3"]), b3=d(p["b3"]), ln_w=d(p["ln_w"]), ln_b=d(p["ln_b"]), has_ln=True, act=0)
E = 5000
e = torch.randn(E, 128, device=dev)
enc = ((torch.randint(0, world, (E,)) << PEER_SHIFT) | torch.randint(0, 1000, (E,))).to(torch.int32).to(dev)
out, _ = ops.mlp_forward([Seg(e), Seg(bufs.local[0], _lib.SEG_GATHER, (enc,))], w, E, _lib.PREC_BF16X3, peer=(bufs.views[0], PEER_SHIFT))
torch.cuda.synchronize()
print(rank, "mlp with peer gather ok", float(out.abs().mean()), flush=True)
dist.barrier(); dist.destroy_process_group()
