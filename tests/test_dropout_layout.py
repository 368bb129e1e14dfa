"""config.training.dropout_rate > 0 (reference Model.py:29-33): a Dropout follows each SiLU, so the Linear layers sit at
Sequential indices 0, 3, 6 instead of 0, 2, 4.  The same layout is built here (checkpoints load); eval-mode forwards are
the dropout-free computation; a training-mode forward drops hidden units (tests/test_gpu_dropout.py)."""
import re
from types import SimpleNamespace as NS

import pytest
import torch

from helpers import LOSS_W, build_model, golden_graphs, mse


def _dropout_model(name, rate=0.1):
    from gnn_fluid_dynamics_b200.models import MODEL_CLASSES
    from fixtures import stats_for
    cfg = NS(model=NS(hidden_width=128, mp_num=15, precision=None, bundle_size=3),
             training=NS(dropout_rate=rate, loss_weights=dict(LOSS_W)))
    return MODEL_CLASSES[name](cfg, mse, None, stats_for(name))


def test_build_mlp_layout_with_dropout():
    """Module order of Model.py:26-35: Linear, SiLU, Dropout, Linear, SiLU, Dropout, Linear (+ LayerNorm wrapper)."""
    from gnn_fluid_dynamics_b200.models.base import build_mlp
    cfg = NS(training=NS(dropout_rate=0.25))
    m = build_mlp(cfg, 10, 128, 128)
    assert [type(x).__name__ for x in m[0]] == ["Linear", "SiLU", "Dropout", "Linear", "SiLU", "Dropout", "Linear"]
    assert all(x.p == 0.25 for x in m[0] if isinstance(x, torch.nn.Dropout))
    assert list(m.state_dict()) == ["0.0.weight", "0.0.bias", "0.3.weight", "0.3.bias", "0.6.weight", "0.6.bias",
                                    "1.weight", "1.bias"]
    assert list(build_mlp(cfg, 10, 128, 3, norm_layer=False).state_dict()) == [
        "0.weight", "0.bias", "3.weight", "3.bias", "6.weight", "6.bias"]
    assert list(build_mlp(NS(training=NS(dropout_rate=0.0)), 10, 128, 128).state_dict()) == [
        "0.0.weight", "0.0.bias", "0.2.weight", "0.2.bias", "0.4.weight", "0.4.bias", "1.weight", "1.bias"]


@pytest.mark.parametrize("name", ["FvgnA", "MgnA", "ConservativeA"])
def test_state_dict_layout_with_dropout(name):
    plain, drop = build_model(name).state_dict(), _dropout_model(name).state_dict()
    assert len(plain) == len(drop)
    assert sorted(tuple(v.shape) for v in plain.values()) == sorted(tuple(v.shape) for v in drop.values())
    assert any(re.search(r"\.[36]\.(weight|bias)$", k) for k in drop), "no Linear at index 3 / 6"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["FvgnA", "MgnA"])
def test_eval_forward_ignores_dropout_and_training_applies_it(name):
    dev = torch.device("cuda:0")
    plain = build_model(name).to(dev).eval()
    drop = _dropout_model(name).to(dev).eval()
    sd, tgt = plain.state_dict(), drop.state_dict()
    src_keys = {re.sub(r"\.3\.(weight|bias)$", r".2.\1", re.sub(r"\.6\.(weight|bias)$", r".4.\1", k)): k for k in tgt}
    drop.load_state_dict({src_keys[k]: v for k, v in sd.items() if k in src_keys} |
                         {k: v for k, v in sd.items() if k in tgt and k not in src_keys.values()}, strict=False)
    for k, dk in src_keys.items():
        assert torch.equal(drop.state_dict()[dk], sd[k]), (k, dk)
    _, graphs = golden_graphs(name, n_cells=300)
    with torch.no_grad():
        a = plain([g.clone().to(dev) for g in graphs], mode="rollout")
        b = drop([g.clone().to(dev) for g in graphs], mode="rollout")
    for k in a:
        assert torch.equal(a[k], b[k]), k
    drop.train()
    with torch.no_grad():
        c = drop([g.clone().to(dev) for g in graphs], mode="rollout")
    assert any(not torch.equal(a[k], c[k]) for k in a) and all(torch.isfinite(c[k]).all() for k in c)


def test_weights_view_folds_the_keep_scale_in_training_only():
    """processor.weights_of (host logic, no GPU): in train() mode a dropout MLP's view carries drop_p and W2, W3 divided by
    (1 - p) - the kernel only masks, the rescale of the kept units lives in the consuming layer's weights - while W1, the
    biases and the LayerNorm affine are untouched; eval() gives the plain view; the cached view follows mode switches and
    optimizer-style in-place updates."""
    from gnn_fluid_dynamics_b200.models.base import build_mlp
    from gnn_fluid_dynamics_b200.processor import weights_of
    torch.manual_seed(0)
    m = build_mlp(NS(training=NS(dropout_rate=0.2)), 12, 128, 128)
    lin = [x for x in m[0] if isinstance(x, torch.nn.Linear)]
    m.train()
    w = weights_of(m)
    assert w.drop_p == pytest.approx(0.2)
    assert torch.equal(w.w1, lin[0].weight) and torch.equal(w.b2, lin[1].bias) and torch.equal(w.ln_w, m[1].weight)
    assert torch.allclose(w.w2 * 0.8, lin[1].weight, rtol=1e-6, atol=0) and torch.allclose(w.w3 * 0.8, lin[2].weight, rtol=1e-6, atol=0)
    assert weights_of(m) is w                                   # cached
    m.eval()
    we = weights_of(m)
    assert we.drop_p == 0.0 and torch.equal(we.w2, lin[1].weight) and torch.equal(we.w3, lin[2].weight)
    m.train()
    with torch.no_grad():
        lin[1].weight.mul_(2.0)                                 # what an optimizer step does
    w2 = weights_of(m)
    assert w2 is not w and torch.allclose(w2.w2 * 0.8, lin[1].weight, rtol=1e-6, atol=0)
    # no dropout configured: train() and eval() share the plain view
    p = build_mlp(NS(training=NS(dropout_rate=0.0)), 12, 128, 128).train()
    assert weights_of(p).drop_p == 0.0 and torch.equal(weights_of(p).w2, [x for x in p[0] if isinstance(x, torch.nn.Linear)][1].weight)


def test_unsupported_dropout_layouts_are_rejected():
    from gnn_fluid_dynamics_b200.models.base import build_mlp
    from gnn_fluid_dynamics_b200.processor import weights_of
    m = build_mlp(NS(training=NS(dropout_rate=0.2)), 12, 128, 128).train()
    drops = [x for x in m[0] if isinstance(x, torch.nn.Dropout)]
    drops[1].p = 0.3                                            # two different rates: not the reference's layout
    with pytest.raises(NotImplementedError, match="dropout"):
        weights_of(m)
    drops[1].p = 0.2
    assert weights_of(m).drop_p == pytest.approx(0.2)


def test_oracle_dropout_forward_is_pinned_to_nn_dropout():
    """oracle.mlp3_dropout (masks as an input) against the reference-layout module itself in train() mode: the masks torch
    drew are read off the Dropout layers with forward hooks (output != 0 wherever the input was non-zero) and plugged into
    the oracle - same output, same parameter gradients (CPU, fp32)."""
    from gnn_fluid_dynamics_b200.models.base import build_mlp
    from oracle.mlp import mlp3_dropout
    torch.manual_seed(3)
    p = 0.3
    m = build_mlp(NS(training=NS(dropout_rate=p)), 16, 128, 128).train()
    x = torch.randn(200, 16)
    masks = []
    hooks = [d.register_forward_hook(lambda mod, inp, out: masks.append((out != 0) | (inp[0] == 0)))
             for d in m[0] if isinstance(d, torch.nn.Dropout)]
    y_ref = m(x)
    for h in hooks:
        h.remove()
    assert len(masks) == 2 and 0.6 < masks[0].float().mean() < 0.8
    lin = [l for l in m[0] if isinstance(l, torch.nn.Linear)]
    y = mlp3_dropout(x, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias,
                     masks[0].float(), masks[1].float(), p, m[1].weight, m[1].bias)
    assert torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6)
    g = torch.randn_like(y)
    ref_grads = torch.autograd.grad((y_ref * g).sum(), list(m.parameters()))
    grads = torch.autograd.grad((y * g).sum(), list(m.parameters()))
    for a, b in zip(grads, ref_grads):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
