"""CPU oracle for the message-passing hot path (TEST INFRASTRUCTURE - never shipped, never measured
as the product).

What this is
------------
A plain CPU/fp32 restatement (torch CPU ops + numpy for the integer work) of the reference's
encoder -> 15 x GN_Block -> decoder path, each function citing the reference file:line it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product package
(``gnn_fluid_dynamics_b200``) never imports ``oracle`` and raises if its CUDA library is missing.

How it is pinned
----------------
The reference ships no tests, golden vectors or fixtures, and it is pure
Python, so the pin is "outputs of the reference itself run here": ``tests/golden/make_golden.py``
imports the reference's own model classes from /root/reference/src (under the 4-module import shim
``tests/golden/refstub.py``), runs them on seeded synthetic meshes with deterministic parameters and
commits the outputs as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures (CPU, no GPU needed).

Third-party arithmetic outside /root/reference that the path relies on and that is restated here:
``torch_scatter.scatter_add`` 2.1.2 (dim=0: zero-initialised output with ``dim_size`` rows, duplicate
indices accumulate, CPU order = ascending source position) and ``torch_geometric`` 2.6.1 ``Data``
attribute semantics.
"""
from .scatter import scatter_add, scatter_add_loop, csr_build  # noqa: F401
from .mlp import mlp3, mlp_from_state  # noqa: F401
from .blocks import (encoder_fwd, gn_block_fwd, decoder_fwd, processor_fwd,  # noqa: F401
                     FAMILIES, family_of)
