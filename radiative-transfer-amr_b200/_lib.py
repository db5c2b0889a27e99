"""ctypes binding of ``csrc/librtb200.so`` (C ABI declared in ``include/rtb200.h``).

Thin by design: argument/return types, the two parameter structs and an error
check.  There is no fallback: if the shared library is missing or was built for
another ABI version, importing the operator fails with a clear message.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 7
MAX_NXY = 12
MAX_NTH = 16

TERM_MASS, TERM_SCAT, TERM_ICONV, TERM_BCONV, TERM_ALL = 1, 2, 4, 8, 15
PC_NONE, PC_NODE, PC_SWEEP, PC_ELEMENT = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librtb200.so")
# kernel-tuning builds (`make lab`: one circulant instantiation, compiles in a minute) are
# loaded only on request; the product and the tests always use librtb200.so
if os.environ.get("RTB200_LAB_LIB") == "1":
    LIB_PATH = os.path.join(_HERE, "csrc", "librtb200_lab.so")

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


class FlatMeshC(C.Structure):
    """Mirror of ``struct rtb200_flatmesh`` (field order matters)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("n_col", C.c_int32), ("n_sig", C.c_int32),
        ("n_sigcell", C.c_int32),
        ("n_dof", C.c_int64), ("n_slab", C.c_int64), ("n_ghost", C.c_int64),
        ("col_nx", _i32p), ("col_ny", _i32p), ("col_sig", _i32p), ("col_bdry", _i32p),
        ("col_dof_off", _i64p), ("col_slab_off", _i64p), ("col_coef_off", _i64p),
        ("col_dx", _f64p), ("col_dy", _f64p),
        ("face_base", _i64p), ("face_rstride", _i32p), ("face_exy", _i32p),
        ("face_n1", _i32p), ("face_acpl", _i32p),
        ("sig_nt", _i32p), ("sig_ntp", _i32p), ("sig_ncell", _i32p), ("sig_cell0", _i32p),
        ("sig_th0", _i64p), ("sig_g_off", _i64p),
        ("cell_nth", _i32p), ("cell_thoff", _i32p),
        ("n_theta_pool", C.c_int64),
        ("th_wt", _f64p), ("th_ct", _f64p), ("th_st", _f64p), ("th_wF", _f64p),
        ("th_cell", _i32p), ("th_bits", _i32p),
        ("n_g_pool", C.c_int64), ("g_pool", _f64p),
        ("n_exy_pool", C.c_int64), ("exy_pool", _f64p),
        ("n_eth_pool", C.c_int64), ("eth_pool", _f64p),
        ("n_acpl_pool", C.c_int64), ("acpl_pool", _i32p),
        ("lgl_w", _f64p), ("lgl_s", _f64p),
        ("n_coef", C.c_int64), ("coef", _f64p),
        ("n_peer", C.c_int32), ("peer_rank", _i32p), ("peer_send0", _i64p),
        ("peer_recv0", _i64p),
        ("n_send", C.c_int64), ("send_src", _i64p),
        ("sig_circ_nc", _i32p), ("sig_circ_p", _i32p), ("sig_circ_off", _i64p),
        ("n_circ_pool", C.c_int64), ("circ_pool", _f64p),
        ("face_nbr", _i32p),
    ]


class GmresParams(C.Structure):
    _fields_ = [("restart", C.c_int32), ("max_it", C.c_int32), ("rtol", C.c_double),
                ("atol", C.c_double), ("divtol", C.c_double), ("orth", C.c_int32),
                ("pc", C.c_int32), ("x0_nonzero", C.c_int32), ("record_history", C.c_int32),
                ("pc_side", C.c_int32)]


class GmresResult(C.Structure):
    _fields_ = [("reason", C.c_int32), ("iterations", C.c_int32), ("res_norm", C.c_double),
                ("res0_norm", C.c_double), ("b_norm", C.c_double),
                ("true_res_norm", C.c_double), ("sigma_max", C.c_double),
                ("sigma_min", C.c_double), ("n_apply", C.c_int64)]


# every symbol include/rtb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "rtb200_abi_version": (C.c_int, []),
    "rtb200_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rtb200_destroy": (None, [C.c_void_p]),
    "rtb200_last_error": (C.c_char_p, [C.c_void_p]),
    "rtb200_set_mesh": (C.c_int, [C.c_void_p, C.POINTER(FlatMeshC)]),
    "rtb200_set_coeffs": (C.c_int, [C.c_void_p, _f64p, C.c_int64]),
    "rtb200_to_slab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtb200_from_slab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtb200_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "rtb200_apply": (C.c_int, [C.c_void_p, C.c_uint, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "rtb200_gmres_work_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "rtb200_gmres": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GmresParams),
                               C.POINTER(GmresResult), _f64p, C.c_void_p, C.c_int64,
                               C.c_void_p]),
    "rtb200_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "rtb200_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "rtb200_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtb200_pc_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtb200_pc_set_shift": (C.c_int, [C.c_void_p, C.c_double]),
    "rtb200_pc_set_eig": (C.c_int, [C.c_void_p, _i64p, C.c_int32, _f64p, C.c_int64]),
    "rtb200_pc_set_fd": (C.c_int, [C.c_void_p, _i64p, C.c_int32, _f64p, C.c_int64]),
    "rtb200_pc_set_type": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rtb200_pc_sweep_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "rtb200_err_angular_jump": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8 + [C.c_int32] * 2
                                + [C.c_void_p] * 4),
    "rtb200_err_spatial_jump": (C.c_int, [C.c_void_p] * 9),
    "rtb200_err_high_resolution": (C.c_int, [C.c_void_p] * 11 + [C.c_int32] + [C.c_void_p] * 3),
    "rtb200_launch_count": (C.c_int64, [C.c_void_p]),
    "rtb200_apply_path": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
}

_lib = None


def load():
    """Load librtb200.so once; raise if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "rt_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C radiative-transfer-amr_b200/csrc` (there is no CPU "
            "fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.rtb200_abi_version() != ABI_VERSION:
        raise RuntimeError("rt_b200: librtb200.so ABI %d, bindings expect %d"
                           % (lib.rtb200_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


class RTB200Error(RuntimeError):
    pass


def check(lib, ctx, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.rtb200_last_error(ctx) if ctx else b""
        raise RTB200Error("%s failed with code %d: %s" % (what, rc, (msg or b"").decode()))
