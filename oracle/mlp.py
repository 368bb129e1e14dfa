"""Oracle: the reference's 3-Linear MLP (Model.py:12-40 / Mgn.py:24-34) and its bias-free Tanh
variant (Conservative.py:31-43), on CPU fp32."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mlp3(x, w1, b1, w2, b2, w3, b3, ln_w=None, ln_b=None, act="silu", ln=None, eps=1e-5):
    """``LN(W3 . act(W2 . act(W1 . x + b1) + b2) + b3)``; weights in PyTorch [out, in] layout.

    act: 'silu' (build_mlp) or 'tanh' (build_mlp_antisym); biases may be None;
    ln=None -> LayerNorm applied iff ln_w is given (eps 1e-5, affine; Model.py:39).
    """
    f = F.silu if act == "silu" else torch.tanh
    h = f(F.linear(x, w1, b1))
    h = f(F.linear(h, w2, b2))
    y = F.linear(h, w3, b3)
    if ln is None:
        ln = ln_w is not None
    if ln:
        y = F.layer_norm(y, (y.shape[-1],), ln_w, ln_b, eps)
    return y


def mlp3_dropout(x, w1, b1, w2, b2, w3, b3, keep1, keep2, p, ln_w=None, ln_b=None, eps=1e-5):
    """Training-mode forward of ``build_mlp`` with ``config.training.dropout_rate = p > 0`` (Model.py:26-35:
    Linear, SiLU, Dropout, Linear, SiLU, Dropout, Linear [, LayerNorm]) for GIVEN keep masks: ``nn.Dropout`` zeroes the
    dropped units and scales the kept ones by 1 / (1 - p).  The masks the reference would draw come from torch's generator
    and cannot be reproduced by another implementation, so the oracle takes them as an input (the CUDA path reports the
    mask it applied through its stash; tests/test_gpu_dropout.py)."""
    s = 1.0 / (1.0 - p)
    h = F.silu(F.linear(x, w1, b1)) * keep1 * s
    h = F.silu(F.linear(h, w2, b2)) * keep2 * s
    y = F.linear(h, w3, b3)
    if ln_w is not None:
        y = F.layer_norm(y, (y.shape[-1],), ln_w, ln_b, eps)
    return y


def mlp_from_state(sd, prefix: str, x, act="silu"):
    """Run the MLP stored under ``prefix`` of a reference-layout state_dict.

    With LayerNorm the module is ``Sequential(Sequential(L,a,L,a,L), LayerNorm)`` so the keys are
    ``prefix.0.{0,2,4}.{weight,bias}`` + ``prefix.1.{weight,bias}``; without, ``prefix.{0,2,4}.*``.
    """
    if f"{prefix}.0.0.weight" in sd:
        p = prefix + ".0"
        ln_w, ln_b = sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"]
    else:
        p = prefix
        ln_w = ln_b = None
    g = lambda k: sd.get(k)
    return mlp3(x, sd[f"{p}.0.weight"], g(f"{p}.0.bias"), sd[f"{p}.2.weight"], g(f"{p}.2.bias"),
                sd[f"{p}.4.weight"], g(f"{p}.4.bias"), ln_w, ln_b, act=act)
