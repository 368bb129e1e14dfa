"""Training-mode dropout of the fused MLP (config.training.dropout_rate > 0, reference src/models/Model.py:29-33: a
Dropout after each SiLU).  The reference draws its masks from torch's generator, which no other implementation can
reproduce, so parity is checked the way a mask-based op allows:

  * the mask the kernel applied is read back from the stash (a dropped unit's saved pre-activation is GNNFD_DROPPED) and
    must equal the documented counter-based hash (include/gnnfd_b200.h: dropout_p), restated here in numpy;
  * with THAT mask plugged into the oracle (oracle.mlp.mlp3_dropout, pinned to nn.Dropout on the CPU), evaluated in fp64,
    the kernel's output, every parameter gradient and every input gradient agree within the fp32-parity tolerance;
  * keep rate, run-to-run behaviour under torch.manual_seed and eval-mode identity at the model level.
"""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from helpers import LOSS_W, golden_graphs, mse
from test_gpu_parity import _rand_mlp, _to_weights, rel_l2

pytestmark = pytest.mark.gpu
DROPPED = -1e30
TOL = 1e-3


def dev():
    return torch.device("cuda:0")


# ---- numpy restatement of csrc/common.cuh: hash32 / dropout_layer_key / dropout_row_hash / dropout_hash ----------------
def _hash32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def expected_dropped(seed, layer, rows, p):
    key = _hash32((seed & 0xFFFFFFFF) ^ int(_hash32(((seed >> 32) + 0x9E3779B9 * (layer + 1)) & 0xFFFFFFFF)))
    r = np.arange(rows, dtype=np.uint64)
    rh = _hash32(((r * 0x9E3779B1) & 0xFFFFFFFF) ^ key)
    h = _hash32((rh[:, None] + np.arange(128, dtype=np.uint64)[None, :]) & 0xFFFFFFFF)
    thresh = int(float(np.float32(p)) * 4294967296.0)
    return h < thresh


def _torch_reference(p, x, m1, m2, drop_p, ln):
    """The oracle's Linear-SiLU-Dropout(mask m1)-Linear-SiLU-Dropout(mask m2)-Linear(-LayerNorm) (oracle.mlp.mlp3_dropout,
    pinned to nn.Dropout on the CPU in tests/test_dropout_layout.py), evaluated in fp64; returns (out, leaf tensors, input)."""
    from oracle.mlp import mlp3_dropout
    leaf = {k: v.double().to(dev()).requires_grad_(True) for k, v in p.items() if v is not None}
    xin = x.double().requires_grad_(True)
    out = mlp3_dropout(xin, leaf["w1"], leaf["b1"], leaf["w2"], leaf["b2"], leaf["w3"], leaf["b3"], m1, m2, drop_p,
                       leaf.get("ln_w") if ln else None, leaf.get("ln_b") if ln else None)
    return out, leaf, xin


def _dropout_weights(p, drop_p):
    """What processor.weights_of builds for a train()-mode module: w2 / w3 pre-divided by (1 - p)."""
    q = dict(p)
    q["w2"], q["w3"] = p["w2"] / (1.0 - drop_p), p["w3"] / (1.0 - drop_p)
    w = _to_weights(q, 0)
    w.drop_p = drop_p
    return w


@pytest.mark.parametrize("rows,drop_p", [(1, 0.5), (130, 0.25), (5000, 0.1)])
def test_mask_is_the_documented_hash(rows, drop_p, monkeypatch):
    from gnn_fluid_dynamics_b200 import ops, _lib
    seed = 0x1234_5678_9ABC_DEF1 & (2 ** 63 - 1)
    monkeypatch.setattr(ops, "dropout_seed", lambda: seed)
    p = _rand_mlp(128, 128, True, seed=3)
    x = torch.randn(rows, 128, generator=torch.Generator().manual_seed(1)).to(dev())
    _, _, st = ops.mlp_forward([ops.Seg(x)], _dropout_weights(p, drop_p), rows, _lib.PREC_BF16X3, stash=True)
    for layer, a in enumerate((st.a1, st.a2)):
        got = (a == DROPPED).cpu().numpy()
        assert np.array_equal(got, expected_dropped(seed, layer, rows, drop_p)), layer
        assert not torch.isnan(a).any()
    # another seed, another mask
    monkeypatch.setattr(ops, "dropout_seed", lambda: seed + 1)
    _, _, st2 = ops.mlp_forward([ops.Seg(x)], _dropout_weights(p, drop_p), rows, _lib.PREC_BF16X3, stash=True)
    if rows > 1:
        assert not torch.equal(st2.a1 == DROPPED, st.a1 == DROPPED)


@pytest.mark.parametrize("prec_name", ["bf16x3", "fp16x3"])
@pytest.mark.parametrize("ln", [True, False])
def test_forward_and_backward_match_torch_with_the_same_mask(prec_name, ln):
    from gnn_fluid_dynamics_b200 import ops, _lib
    from gnn_fluid_dynamics_b200.precisions import available
    if prec_name not in available():
        pytest.skip(prec_name)
    prec = _lib.PRECISIONS[prec_name]
    rows, drop_p = 3000, 0.2
    p = _rand_mlp(128, 128, ln, seed=5)
    g0 = torch.Generator().manual_seed(2)
    x = torch.randn(rows, 128, generator=g0).to(dev())
    gout = torch.randn(rows, 128, generator=g0).to(dev())
    w = _dropout_weights(p, drop_p)
    raw, _, st = ops.mlp_forward([ops.Seg(x)], w, rows, prec, stash=True)
    m1, m2 = (st.a1 != DROPPED).double(), (st.a2 != DROPPED).double()
    assert abs(float(m1.mean()) - (1 - drop_p)) < 0.01 and abs(float(m2.mean()) - (1 - drop_p)) < 0.01
    ref, leaf, xin = _torch_reference(p, x, m1, m2, drop_p, ln)
    assert rel_l2(raw, ref.float()) < TOL
    ref.backward(gout.double())
    ws = ops.mlp_backward_workspace(rows, dev())
    grads, dins = ops.mlp_backward([ops.Seg(x)], w, st, rows, gout, prec, [{}], ws)
    for k, t in leaf.items():
        assert rel_l2(grads[k], t.grad.float()) < TOL, (k, rel_l2(grads[k], t.grad.float()))
    assert rel_l2(dins[0], xin.grad.float()) < TOL


def test_stepwise_backward_equals_fused_call_under_dropout():
    from gnn_fluid_dynamics_b200 import ops, _lib, training
    rows, drop_p = 1000, 0.3
    p = _rand_mlp(128, 128, True, seed=7)
    x = torch.randn(rows, 128, generator=torch.Generator().manual_seed(4)).to(dev())
    gout = torch.randn(rows, 128, generator=torch.Generator().manual_seed(5)).to(dev())
    w = _dropout_weights(p, drop_p)
    _, _, st = ops.mlp_forward([ops.Seg(x)], w, rows, _lib.PREC_BF16X3, stash=True)
    ws = ops.mlp_backward_workspace(rows, dev())
    ga, da = training.mlp_backward(w, st, [ops.Seg(x)], rows, gout, _lib.PREC_BF16X3, [{}], ws)
    gb, db = training.mlp_backward_stepwise(w, st, [ops.Seg(x)], rows, gout, _lib.PREC_BF16X3, [{}], ws)
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None:
            assert rel_l2(a, b) < 1e-5
    assert rel_l2(da[0], db[0]) < 1e-5


def _dropout_model(name, rate):
    from gnn_fluid_dynamics_b200.models import MODEL_CLASSES
    from fixtures import stats_for
    cfg = NS(model=NS(hidden_width=128, mp_num=15, precision=None, bundle_size=3),
             training=NS(dropout_rate=rate, loss_weights=dict(LOSS_W)))
    return MODEL_CLASSES[name](cfg, mse, None, stats_for(name))


def _step(model, graphs):
    model.zero_grad(set_to_none=True)
    out = model([g.clone().to(dev()) for g in graphs], mode="train")
    gn = model.normalizer.input([g.clone().to(dev()) for g in graphs])
    loss = model.loss(out, gn)["total_log_loss"]
    loss.backward()
    return loss.detach().clone(), [p.grad.clone() for p in model.parameters() if p.grad is not None]


@pytest.mark.parametrize("name", ["FvgnA", "MgnA", "ConservativeA"])
def test_model_trains_with_dropout(name):
    """Hand-scheduled backward (Fvgn / Mgn) and the per-op autograd path (Conservative): a train()-mode step with
    dropout is finite, differs from the dropout-free step, repeats under torch.manual_seed and changes with the seed."""
    torch.manual_seed(0)
    model = _dropout_model(name, 0.1).to(dev()).train()
    _, graphs = golden_graphs(name, flip=True, n_cells=1500)
    torch.manual_seed(11)
    l1, g1 = _step(model, graphs)
    torch.manual_seed(11)
    l2, g2 = _step(model, graphs)
    torch.manual_seed(12)
    l3, g3 = _step(model, graphs)
    assert torch.isfinite(l1) and all(torch.isfinite(g).all() for g in g1) and len(g1) > 0
    assert torch.equal(l1, l2) and all(torch.equal(a, b) for a, b in zip(g1, g2))
    assert not torch.equal(l1, l3)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    l0, _ = _step(model, graphs)
    assert not torch.equal(l0, l1)


def test_train_mode_forward_without_grad_also_drops():
    """model.train() + torch.no_grad() (a validation pass that forgot eval()): the reference drops units there too."""
    model = _dropout_model("FvgnA", 0.2).to(dev())
    _, graphs = golden_graphs("FvgnA", n_cells=600)
    with torch.no_grad():
        model.eval()
        a = model([g.clone().to(dev()) for g in graphs], mode="rollout")
        b = model([g.clone().to(dev()) for g in graphs], mode="rollout")
        model.train()
        c = model([g.clone().to(dev()) for g in graphs], mode="rollout")
    k = next(iter(a))
    assert torch.equal(a[k], b[k]) and not torch.equal(a[k], c[k]) and torch.isfinite(c[k]).all()


def test_bad_dropout_arguments_are_rejected():
    from gnn_fluid_dynamics_b200 import ops, _lib
    p = _rand_mlp(128, 128, True, seed=3)
    x = torch.randn(10, 128).to(dev())
    w = _dropout_weights(p, 0.5)
    with pytest.raises(RuntimeError, match="tensor-core precision"):
        ops.mlp_forward([ops.Seg(x)], w, 10, _lib.PREC_F32)
    w.drop_p = 1.0
    with pytest.raises(RuntimeError, match=r"\[0, 1\)"):
        ops.mlp_forward([ops.Seg(x)], w, 10, _lib.PREC_BF16X3)
