"""ctypes binding of the C-ABI library ``lib/libgnnfd_b200.so`` (declared in include/gnnfd_b200.h).

There is deliberately no fallback: if the library is missing or an entry point is absent the import
fails loudly, and every op raises ``RuntimeError`` on a non-zero status.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C gnn_fluid_dynamics_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GNNFD_LIB: experiment builds of the same library (scripts/abl_edge.py); the default is the in-tree build
LIB_PATH = os.environ.get("GNNFD_LIB") or os.path.join(_HERE, "lib", "libgnnfd_b200.so")

PREC_F32, PREC_BF16X3, PREC_BF16X1, PREC_FP16X2, PREC_FP16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"f32": PREC_F32, "bf16x3": PREC_BF16X3, "bf16x1": PREC_BF16X1,
              "fp16x2": PREC_FP16X2, "fp16x3": PREC_FP16X3}
ACT_SILU, ACT_TANH = 0, 1
SEG_DIRECT, SEG_GATHER, SEG_SUM2, SEG_DIFF2, SEG_MEAN3, SEG_SUM3S = 0, 1, 2, 3, 4, 5
SUM3S_ZERO = -(2 ** 31)          # GNNFD_SUM3S_ZERO: index entry that contributes nothing

EXPORTS = [
    "gnnfd_abi_version", "gnnfd_last_error", "gnnfd_index_narrow", "gnnfd_csr_workspace_bytes",
    "gnnfd_csr_build", "gnnfd_segment_sum", "gnnfd_mlp_forward", "gnnfd_pack_mlp_bytes",
    "gnnfd_pack_mlp", "gnnfd_tc_profile_read",
    "gnnfd_ln_backward_workspace_bytes", "gnnfd_ln_backward", "gnnfd_wgrad_workspace_bytes", "gnnfd_wgrad",
    "gnnfd_segment_sum3", "gnnfd_gather_pair_add", "gnnfd_struct_size",
    "gnnfd_mlp_backward_workspace_bytes", "gnnfd_pack_mlp_backward_bytes", "gnnfd_pack_mlp_backward",
    "gnnfd_mlp_backward", "gnnfd_gather_rows", "gnnfd_enable_peer_access", "gnnfd_gather_cols_add",
    "gnnfd_glue_workspace_bytes", "gnnfd_face_area_norm", "gnnfd_face_area_norm_backward", "gnnfd_fvm_integrate",
    "gnnfd_fvm_integrate_backward", "gnnfd_masked_mse", "gnnfd_masked_mse_backward", "gnnfd_state_advance",
    "gnnfd_affine_columns", "gnnfd_set_launch_overlap", "gnnfd_set_l2_hints", "gnnfd_flux_integrate", "gnnfd_gather3", "gnnfd_gather3_backward", "gnnfd_dropout_hash",
]
ABI_VERSION = 5


class Segment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("idx", C.c_void_p * 3), ("ld", C.c_int32), ("col", C.c_int32),
                ("width", C.c_int32), ("mode", C.c_int32), ("split", C.c_void_p), ("src_rows", C.c_int64)]


class MlpArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("n_seg", C.c_int32), ("seg", Segment * 3),
        ("k_in", C.c_int32), ("hidden", C.c_int32), ("n_out", C.c_int32),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
        ("has_ln", C.c_int32), ("ln_eps", C.c_float), ("act", C.c_int32),
        ("mul", C.c_void_p), ("residual", C.c_void_p), ("out_raw", C.c_void_p),
        ("out_sum", C.c_void_p), ("packed", C.c_void_p), ("precision", C.c_int32),
        ("n_layers", C.c_int32), ("mul_mode", C.c_int32),
        ("save_a1", C.c_void_p), ("save_a2", C.c_void_p), ("save_rstd", C.c_void_p), ("save_xhat", C.c_void_p),
        ("w1_ld_n", C.c_int32), ("w1_ld_k", C.c_int32), ("w1_rows", C.c_int32),
        ("bwd_chain", C.c_int32), ("hid_mul1", C.c_void_p), ("hid_mul2", C.c_void_p),
        ("w2_ld_n", C.c_int32), ("w2_ld_k", C.c_int32), ("w3_ld_n", C.c_int32), ("w3_ld_k", C.c_int32),
        ("w3_rows", C.c_int32),
        ("peer_base", C.c_void_p * 8), ("peer_shift", C.c_int32),
        ("out_split", C.c_void_p), ("split_of_sum", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("static_operands", C.c_int32),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("a", Segment), ("a_act", C.c_int32), ("n_b", C.c_int32), ("b", Segment * 3),
        ("b_act", C.c_int32), ("out", C.c_void_p), ("ld_out", C.c_int32), ("transpose_out", C.c_int32),
        ("colsum", C.c_void_p), ("colsum_of_b", C.c_int32), ("precision", C.c_int32),
    ]


class MlpBackwardArgs(C.Structure):
    _fields_ = [
        ("fwd", MlpArgs), ("g", C.c_void_p),
        ("a1", C.c_void_p), ("a2", C.c_void_p), ("xhat", C.c_void_p), ("rstd", C.c_void_p),
        ("packed_bwd", C.c_void_p),
        ("d_w1", C.c_void_p), ("d_b1", C.c_void_p), ("d_w2", C.c_void_p), ("d_b2", C.c_void_p),
        ("d_w3", C.c_void_p), ("d_b3", C.c_void_p), ("d_ln_w", C.c_void_p), ("d_ln_b", C.c_void_p),
        ("din_out", C.c_void_p * 3), ("din_residual", C.c_void_p * 3),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("da1_out", C.c_void_p), ("skip_wgrad_l1", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library of gnn_fluid_dynamics_b200 is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} does not export {missing}")
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.gnnfd_abi_version.restype = C.c_int
    lib.gnnfd_set_launch_overlap.argtypes = [C.c_int32]
    lib.gnnfd_set_l2_hints.argtypes = [C.c_int32]
    lib.gnnfd_dropout_hash.argtypes = [C.c_uint64, C.c_int32, C.c_uint32, C.c_uint32]
    lib.gnnfd_dropout_hash.restype = C.c_uint32
    lib.gnnfd_last_error.restype = C.c_char_p
    lib.gnnfd_index_narrow.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.gnnfd_csr_workspace_bytes.argtypes = [i64, i64]
    lib.gnnfd_csr_workspace_bytes.restype = C.c_size_t
    lib.gnnfd_csr_build.argtypes = [vp, i64, i64, vp, vp, vp, C.c_size_t, vp]
    lib.gnnfd_segment_sum.argtypes = [vp, vp, i32, i32, i32, i32, i32, f32, i64, vp, vp, i64, vp, i32, vp]
    lib.gnnfd_mlp_forward.argtypes = [C.POINTER(MlpArgs), vp]
    lib.gnnfd_pack_mlp_bytes.argtypes = [i32, i32, i32, i32]
    lib.gnnfd_pack_mlp_bytes.restype = C.c_size_t
    lib.gnnfd_pack_mlp.argtypes = [C.POINTER(MlpArgs), vp, vp]
    lib.gnnfd_tc_profile_read.argtypes = [C.POINTER(C.c_uint64)]
    lib.gnnfd_ln_backward_workspace_bytes.argtypes = [i64]
    lib.gnnfd_ln_backward_workspace_bytes.restype = C.c_size_t
    lib.gnnfd_ln_backward.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, C.c_size_t, vp]
    lib.gnnfd_wgrad_workspace_bytes.argtypes = [i64, i32]
    lib.gnnfd_wgrad_workspace_bytes.restype = C.c_size_t
    lib.gnnfd_wgrad.argtypes = [C.POINTER(WgradArgs), vp, C.c_size_t, vp]
    lib.gnnfd_segment_sum3.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, f32, i64, vp, vp, i64, f32, vp, i32,
                                       vp, i32, vp]
    lib.gnnfd_gather_pair_add.argtypes = [vp, vp, vp, i32, vp, vp, f32, i32, i64, vp]
    lib.gnnfd_gather_rows.argtypes = [vp, i32, vp, i64, i32, vp, vp]
    lib.gnnfd_enable_peer_access.argtypes = [i32]
    lib.gnnfd_gather_cols_add.argtypes = [vp, i32, i32, i32, vp, i32, vp, f32, i64, vp]
    lib.gnnfd_struct_size.argtypes = [i32]
    lib.gnnfd_struct_size.restype = C.c_size_t
    lib.gnnfd_mlp_backward_workspace_bytes.argtypes = [C.POINTER(MlpArgs)]
    lib.gnnfd_mlp_backward_workspace_bytes.restype = C.c_size_t
    lib.gnnfd_pack_mlp_backward_bytes.argtypes = [C.POINTER(MlpArgs)]
    lib.gnnfd_pack_mlp_backward_bytes.restype = C.c_size_t
    lib.gnnfd_pack_mlp_backward.argtypes = [C.POINTER(MlpArgs), vp, i32, vp]
    lib.gnnfd_mlp_backward.argtypes = [C.POINTER(MlpBackwardArgs), vp]
    sz = C.c_size_t
    lib.gnnfd_glue_workspace_bytes.argtypes = []
    lib.gnnfd_glue_workspace_bytes.restype = sz
    lib.gnnfd_face_area_norm.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, i32, f32, f32, i32, vp, vp, vp, sz, vp]
    lib.gnnfd_face_area_norm_backward.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, f32, vp, vp, vp, vp, sz, vp]
    lib.gnnfd_gather3.argtypes = [vp, i32, i32, vp, vp, vp, i64, vp, vp]
    lib.gnnfd_gather3_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, i64, vp, i32, vp]
    lib.gnnfd_flux_integrate.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, f32, vp, vp, f32, f32, vp]
    lib.gnnfd_fvm_integrate.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, f32, vp, vp, vp]
    lib.gnnfd_fvm_integrate_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, f32, vp, vp, vp, i32, i32, vp, vp]
    lib.gnnfd_masked_mse.argtypes = [vp, i32, vp, i32, vp, i64, i32, vp, vp, sz, vp]
    lib.gnnfd_masked_mse_backward.argtypes = [vp, i32, vp, i32, vp, i64, i32, vp, vp, vp, i32, vp]
    lib.gnnfd_state_advance.argtypes = [vp, i32, vp, i32, i32, i64, vp, i32, vp, vp, vp, vp, vp, i32, i64, vp, i32, vp, i32,
                                        vp, vp, vp]
    lib.gnnfd_affine_columns.argtypes = [vp, i64, i32, i32, vp, vp, vp, i32, vp]
    for which, mirror in ((0, MlpArgs), (1, WgradArgs), (2, Segment), (3, MlpBackwardArgs)):
        if lib.gnnfd_struct_size(which) != C.sizeof(mirror):
            raise ImportError(f"{LIB_PATH}: struct {mirror.__name__} is {lib.gnnfd_struct_size(which)} bytes in the "
                              f"library but {C.sizeof(mirror)} in the binding; rebuild")
    if lib.gnnfd_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.gnnfd_abi_version()} != {ABI_VERSION}; rebuild")
    return lib


lib = _load()


_launch_overlap = None
# inference on meshes up to this many faces is launch-latency bound: overlap the launches (see include/gnnfd_b200.h)
LAUNCH_OVERLAP_MAX_FACES = 32768


def set_launch_overlap(on: bool) -> None:
    """Launch policy of the library's kernels (gnnfd_set_launch_overlap); cached, so calling it per forward is free."""
    global _launch_overlap
    on = bool(on)
    if on != _launch_overlap:
        lib.gnnfd_set_launch_overlap(int(on))
        _launch_overlap = on


def choose_launch_overlap(n_faces: int, training: bool) -> None:
    """Training steps (hundreds of short dependent launches per step) and launch-bound small meshes overlap their
    launches; large-mesh inference steps measured 1-2 % slower with it and keep plain stream order."""
    set_launch_overlap(training or n_faces <= LAUNCH_OVERLAP_MAX_FACES)


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with status {rc}: {lib.gnnfd_last_error().decode()}")
