"""Which GEMM arithmetic variants the built library implements (see include/gnnfd_b200.h)."""
from __future__ import annotations

from . import _lib


def available():
    """Names accepted by ``Model.set_precision`` that this build supports."""
    out = ["f32"]
    for name, code in _lib.PRECISIONS.items():
        if name != "f32" and _lib.lib.gnnfd_pack_mlp_bytes(384, 128, 128, code) > 0:
            out.append(name)
    return out
