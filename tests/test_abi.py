"""The C-ABI library loads and exports every symbol include/gnnfd_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gnnfd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnnfd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from gnn_fluid_dynamics_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 9
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gnnfd_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "python binding list and header disagree"


def test_abi_version_and_struct_layout():
    from gnn_fluid_dynamics_b200 import _lib
    assert _lib.lib.gnnfd_abi_version() == _lib.ABI_VERSION == 5
    for which, mirror in ((0, _lib.MlpArgs), (1, _lib.WgradArgs), (2, _lib.Segment)):
        assert _lib.lib.gnnfd_struct_size(which) == ctypes.sizeof(mirror)
    # gnnfd_segment: ptr + 3 ptr + 4 int32 + split ptr + int64 src_rows = 64 bytes; args struct must be 8-byte aligned
    assert ctypes.sizeof(_lib.Segment) == 64
    assert ctypes.sizeof(_lib.MlpArgs) % 8 == 0


def test_bad_arguments_return_status_not_crash():
    from gnn_fluid_dynamics_b200 import _lib
    args = _lib.MlpArgs()
    args.rows = -1
    rc = _lib.lib.gnnfd_mlp_forward(ctypes.byref(args), None)
    assert rc == -1
    assert b"rows" in _lib.lib.gnnfd_last_error()
    assert _lib.lib.gnnfd_csr_workspace_bytes(1000, 100) > 4 * 1100


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gnn_fluid_dynamics_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), fn


def test_dropout_hash_known_answers_and_numpy_restatement():
    """The dropout mask function (gnnfd_mlp_args.dropout_p) evaluated by the library on the HOST against (a) the numpy
    restatement the GPU tests compare the kernel's mask with (tests/test_gpu_dropout.py) and (b) its statistics: a fair
    coin per unit, independent across rows, columns, layers and seeds."""
    import importlib.util
    import numpy as np
    from gnn_fluid_dynamics_b200 import _lib
    spec = importlib.util.spec_from_file_location("t_dropout", os.path.join(ROOT, "tests", "test_gpu_dropout.py"))
    src = open(spec.origin).read()
    ns = {"np": np}
    exec(src[src.index("def _hash32"):src.index("def _torch_reference")], ns)      # the restatement only (no CUDA imports)
    seed = 0x0123_4567_89AB_CDEF
    for layer in (0, 1):
        ref = ns["expected_dropped"](seed, layer, 40, 0.25)
        thresh = int(float(np.float32(0.25)) * 4294967296.0)
        got = np.array([[_lib.lib.gnnfd_dropout_hash(seed, layer, r, c) < thresh for c in range(128)] for r in range(40)])
        assert np.array_equal(got, ref), layer
    # statistics of the library's own function: keep rate and independence between layers / neighbouring seeds
    h = lambda s, l: np.array([[_lib.lib.gnnfd_dropout_hash(s, l, r, c) for c in range(128)] for r in range(64)], dtype=np.float64)
    a, b, c = h(seed, 0), h(seed, 1), h(seed + 1, 0)
    for x in (a, b, c):
        assert abs(x.mean() / 2 ** 32 - 0.5) < 0.01
    assert abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 0.05 and abs(np.corrcoef(a.ravel(), c.ravel())[0, 1]) < 0.05
    assert abs(np.corrcoef(a[:, :-1].ravel(), a[:, 1:].ravel())[0, 1]) < 0.05      # neighbouring columns
    assert abs(np.corrcoef(a[:-1].ravel(), a[1:].ravel())[0, 1]) < 0.05            # neighbouring rows
