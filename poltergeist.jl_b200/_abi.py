"""ctypes view of include/poltergeist_b200.h and the loader of libpoltergeist_b200.so.

There is no CPU fallback: if the CUDA library is missing or no device is present the
calls raise (PGB_ERR_NO_DEVICE); nothing in this package routes to the test oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpoltergeist_b200.so")

# ---- enums (keep in sync with include/poltergeist_b200.h) --------------------
PGB_OK = 0
PGB_ERR_BAD_ARG, PGB_ERR_CUDA, PGB_ERR_NEWTON, PGB_ERR_DOMAIN = 1, 2, 3, 4
PGB_ERR_NO_DEVICE, PGB_ERR_SOLVER, PGB_ERR_UNSUPPORTED, PGB_ERR_NOT_CONVERGED = 5, 6, 7, 8

FAM_POLYSINE, FAM_SQRT, FAM_SINASIN, FAM_CHEB, FAM_MOBIUS = 0, 1, 2, 3, 4
FORWARD, REVERSE = 0, 1
CHEBYSHEV, FOURIER = 0, 1
MODE_INTERVAL, MODE_FWD_CIRCLE, MODE_REV_CIRCLE = 0, 1, 2


class pgb_branch(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("dir", C.c_int32), ("mode", C.c_int32),
        ("ncoef", C.c_int32), ("coef_off", C.c_int32),
        ("ndcoef", C.c_int32), ("dcoef_off", C.c_int32), ("reserved", C.c_int32),
        ("dom_a", C.c_double), ("dom_b", C.c_double),
        ("ran_a", C.c_double), ("ran_b", C.c_double),
        ("shift", C.c_double), ("fa", C.c_double), ("fb", C.c_double),
        ("par", C.c_double * 8),
    ]


class pgb_stage(C.Structure):
    _fields_ = [("nbranch", C.c_int32), ("branch_off", C.c_int32)]


class pgb_map_desc(C.Structure):
    _fields_ = [
        ("nstage", C.c_int32), ("nbranch", C.c_int32), ("ncoef", C.c_int32),
        ("domain_space", C.c_int32), ("range_space", C.c_int32), ("reserved", C.c_int32),
        ("dom_a", C.c_double), ("dom_b", C.c_double),
        ("ran_a", C.c_double), ("ran_b", C.c_double),
        ("stages", C.POINTER(pgb_stage)),
        ("branches", C.POINTER(pgb_branch)),
        ("coefs", C.POINTER(C.c_double)),
        ("npath", C.c_int32), ("reserved2", C.c_int32),
        ("path_ptr", C.POINTER(C.c_int32)),
        ("path_branch", C.POINTER(C.c_int32)),
        ("dvmin", C.c_double),
    ]


class pgb_chop_opts(C.Structure):
    _fields_ = [("tol", C.c_double), ("chop", C.c_int32), ("reserved", C.c_int32)]


class PoltergeistError(RuntimeError):
    """Raised for every non-zero status of the C ABI (mirrors the reference's `error(...)`)."""

    def __init__(self, status, text):
        super().__init__(f"pgb status {status}: {text}")
        self.status = status


_lib = None

# every symbol include/poltergeist_b200.h declares (checked by tests/test_abi_symbols.py)
EXPORTS = [
    "pgb_abi_version", "pgb_last_error", "pgb_ctx_create", "pgb_ctx_destroy", "pgb_ctx_sync",
    "pgb_ctx_stream", "pgb_ctx_launch_count", "pgb_ctx_timer_start", "pgb_ctx_timer_stop",
    "pgb_ctx_profile", "pgb_ctx_profile_read", "pgb_ctx_set_stream", "pgb_bench_fp64_peak",
    "pgb_dev_alloc", "pgb_dev_free", "pgb_dev_download", "pgb_dev_upload",
    "pgb_map_create", "pgb_map_destroy", "pgb_map_invalidate", "pgb_map_ncomposite", "pgb_invert_branches",
    "pgb_ipc_export", "pgb_ipc_open", "pgb_ipc_close", "pgb_zero_columns", "pgb_zero_columns_ragged", "pgb_peer_barrier", "pgb_ctx_graph_begin", "pgb_ctx_graph_end", "pgb_graph_launch",
    "pgb_graph_destroy", "pgb_transfer_fill_device_multi",
    "pgb_transfer_fill", "pgb_transfer_fill_device", "pgb_transfer_fill_device_sparse", "pgb_transfer_fill_ragged", "pgb_transfer_fill_ragged_tail", "pgb_transfer_nodes_used",
    "pgb_transfer_create", "pgb_transfer_destroy", "pgb_transfer_resize", "pgb_transfer_set_speculation", "pgb_transfer_ncols",
    "pgb_transfer_colstop", "pgb_transfer_ragged_size", "pgb_transfer_get_ragged",
    "pgb_transfer_get_dense", "pgb_transfer_apply",
    "pgb_solinv_create", "pgb_solinv_create_dense", "pgb_solinv_factored_block", "pgb_solinv_destroy", "pgb_solinv_solve",
    "pgb_solinv_acim", "pgb_eigvals", "pgb_eigvals_dense",
    "pgb_batch_create", "pgb_batch_destroy", "pgb_batch_fill_acim", "pgb_batch_eigvals",
    "pgb_group_create", "pgb_group_destroy", "pgb_group_size", "pgb_group_ctx", "pgb_group_sync", "pgb_group_map_create",
    "pgb_group_map_destroy", "pgb_group_map", "pgb_gmatrix_create", "pgb_gmatrix_destroy", "pgb_gmatrix_ptr", "pgb_gmatrix_colstops",
    "pgb_gmatrix_multicast", "pgb_group_fill", "pgb_group_fill_ragged", "pgb_column_shard", "pgb_host_alloc", "pgb_host_free",
    "pgb_host_register", "pgb_host_unregister", "pgb_symm_begin", "pgb_symm_connect", "pgb_symm_bind", "pgb_symm_ptr",
    "pgb_symm_multicast_ptr", "pgb_symm_size", "pgb_symm_destroy",
]


def lib():
    """Load libpoltergeist_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PoltergeistError(
            PGB_ERR_NO_DEVICE,
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            " There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "pgb_abi_version": (C.c_int, []),
        "pgb_last_error": (C.c_char_p, [vp]),
        "pgb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "pgb_ctx_destroy": (None, [vp]),
        "pgb_ctx_sync": (C.c_int, [vp]),
        "pgb_ctx_stream": (vp, [vp]),
        "pgb_ctx_launch_count": (i64, [vp]),
        "pgb_ctx_set_stream": (C.c_int, [vp, vp]),
        "pgb_dev_alloc": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "pgb_dev_free": (C.c_int, [vp, vp]),
        "pgb_dev_download": (C.c_int, [vp, vp, vp, i64]),
        "pgb_dev_upload": (C.c_int, [vp, vp, vp, i64]),
        "pgb_ctx_timer_start": (C.c_int, [vp]),
        "pgb_ctx_timer_stop": (C.c_int, [vp, dp]),
        "pgb_bench_fp64_peak": (C.c_int, [vp, dp]),
        "pgb_ctx_profile": (C.c_int, [vp, C.c_int]),
        "pgb_ctx_profile_read": (C.c_int, [vp, C.c_int, dp, i64p]),
        "pgb_map_create": (C.c_int, [vp, C.POINTER(pgb_map_desc), C.POINTER(vp)]),
        "pgb_map_destroy": (None, [vp]),
        "pgb_map_invalidate": (C.c_int, [vp]),
        "pgb_map_ncomposite": (i32, [vp]),
        "pgb_invert_branches": (C.c_int, [vp, vp, i64, dp, dp, dp]),
        "pgb_transfer_fill": (C.c_int, [vp, vp, i64, i64, i64, i64, dp, i64, i32p, C.POINTER(pgb_chop_opts)]),
        "pgb_transfer_fill_device": (C.c_int, [vp, vp, i64, i64, i64, i64, vp, i64, vp, C.POINTER(pgb_chop_opts)]),
        "pgb_transfer_fill_device_sparse": (C.c_int, [vp, vp, i64, i64, i64, i64, vp, i64, vp, vp, C.POINTER(pgb_chop_opts)]),
        "pgb_ipc_export": (C.c_int, [vp, vp, vp]),
        "pgb_ipc_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "pgb_ipc_close": (C.c_int, [vp, vp]),
        "pgb_zero_columns": (C.c_int, [vp, vp, i64, i64, i64, i64]),
        "pgb_zero_columns_ragged": (C.c_int, [vp, vp, i64, vp, i64, i64]),
        "pgb_peer_barrier": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.c_int, vp]),
        "pgb_ctx_graph_begin": (C.c_int, [vp]),
        "pgb_ctx_graph_end": (C.c_int, [vp, C.POINTER(vp)]),
        "pgb_graph_launch": (C.c_int, [vp, vp]),
        "pgb_graph_destroy": (None, [vp]),
        "pgb_transfer_fill_device_multi": (C.c_int, [vp, vp, i64, i64, i64, i64, C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int, i64,
                                                    C.POINTER(pgb_chop_opts), C.POINTER(vp), vp, vp, vp, vp, vp, vp, i64]),
        "pgb_transfer_fill_ragged_tail": (C.c_int, [vp, vp, i64, i64, i64, dp, i64, i64p, i64p, i32p, i64p, i64p, C.POINTER(pgb_chop_opts)]),
        "pgb_transfer_fill_ragged": (C.c_int, [vp, vp, i64, i64, i64, dp, i64, i64p, i32p, i64p, i64p, C.POINTER(pgb_chop_opts)]),
        "pgb_transfer_nodes_used": (C.c_int, [vp, i64, i64, i32p]),
        "pgb_transfer_create": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "pgb_transfer_destroy": (None, [vp]),
        "pgb_transfer_resize": (C.c_int, [vp, i64]),
        "pgb_transfer_set_speculation": (C.c_int, [vp, C.c_int]),
        "pgb_transfer_ncols": (i64, [vp]),
        "pgb_transfer_colstop": (C.c_int, [vp, i64, i32p]),
        "pgb_transfer_ragged_size": (C.c_int, [vp, i64, i64, i64p]),
        "pgb_transfer_get_ragged": (C.c_int, [vp, i64, i64, dp, i64, i64p, i64p]),
        "pgb_transfer_get_dense": (C.c_int, [vp, i64, i64, i64, dp, i64]),
        "pgb_transfer_apply": (C.c_int, [vp, dp, i64, dp, i64]),
        "pgb_solinv_create": (C.c_int, [vp, i64, dp, i64, C.POINTER(vp)]),
        "pgb_solinv_create_dense": (C.c_int, [vp, vp, i64, i64, vp, i32, C.c_double, C.c_double, dp, i64, C.POINTER(vp)]),
        "pgb_solinv_factored_block": (i64, [vp]),
        "pgb_solinv_destroy": (None, [vp]),
        "pgb_solinv_solve": (C.c_int, [vp, dp, dp, i64]),
        "pgb_solinv_acim": (C.c_int, [vp, dp]),
        "pgb_eigvals": (C.c_int, [vp, i64, dp, dp]),
        "pgb_eigvals_dense": (C.c_int, [vp, vp, i64, i64, dp, dp]),
        "pgb_batch_create": (C.c_int, [vp, C.POINTER(pgb_map_desc), i32, C.POINTER(vp)]),
        "pgb_batch_destroy": (None, [vp]),
        "pgb_batch_fill_acim": (C.c_int, [vp, i64, i64, dp, dp]),
        "pgb_batch_eigvals": (C.c_int, [vp, i64, i64, i64, dp, dp]),
        "pgb_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "pgb_group_destroy": (None, [vp]),
        "pgb_group_size": (C.c_int, [vp]),
        "pgb_group_ctx": (vp, [vp, C.c_int]),
        "pgb_group_sync": (C.c_int, [vp]),
        "pgb_group_map_create": (C.c_int, [vp, C.POINTER(pgb_map_desc), C.POINTER(vp)]),
        "pgb_group_map_destroy": (None, [vp]),
        "pgb_group_map": (vp, [vp, C.c_int]),
        "pgb_gmatrix_create": (C.c_int, [vp, i64, i64, C.c_int, C.POINTER(vp)]),
        "pgb_gmatrix_destroy": (None, [vp]),
        "pgb_gmatrix_ptr": (vp, [vp, C.c_int]),
        "pgb_gmatrix_colstops": (vp, [vp, C.c_int]),
        "pgb_gmatrix_multicast": (C.c_int, [vp]),
        "pgb_group_fill": (C.c_int, [vp, vp, i64, vp, C.POINTER(pgb_chop_opts)]),
        "pgb_group_fill_ragged": (C.c_int, [vp, vp, i64, i64, i64, dp, i64, i64p, i32p, i64p, i64p, C.POINTER(pgb_chop_opts)]),
        "pgb_column_shard": (C.c_int, [i64, C.c_int, C.c_int, i64p, i64p]),
        "pgb_host_alloc": (C.c_int, [i64, C.POINTER(vp)]),
        "pgb_host_free": (C.c_int, [vp]),
        "pgb_host_register": (C.c_int, [vp, i64]),
        "pgb_host_unregister": (C.c_int, [vp]),
        "pgb_symm_begin": (C.c_int, [vp, C.c_int, C.c_int, i64, C.c_int, C.POINTER(vp), i32p, i32p]),
        "pgb_symm_connect": (C.c_int, [vp, i32p, i32p, i32]),
        "pgb_symm_bind": (C.c_int, [vp]),
        "pgb_symm_ptr": (vp, [vp, C.c_int]),
        "pgb_symm_multicast_ptr": (vp, [vp]),
        "pgb_symm_size": (i64, [vp]),
        "pgb_symm_destroy": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status, ctx=None):
    if status != PGB_OK:
        msg = lib().pgb_last_error(ctx)
        raise PoltergeistError(status, msg.decode() if msg else "")


def dptr(a):
    """double* of a C-contiguous float64 numpy array (or NULL)."""
    if a is None:
        return C.cast(None, C.POINTER(C.c_double))
    return a.ctypes.data_as(C.POINTER(C.c_double))
