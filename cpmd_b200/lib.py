"""ctypes binding of the C ABI in ``include/cpb200.h`` (libcpb200.so, sm_100a).

The product path has no CPU fallback: :func:`load` loads ``cpmd_b200/libcpb200.so`` (built
in-tree by ``__graft_entry__.build()`` / ``make -C cpmd_b200/csrc``) or raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CPB200_LIB: tuning hook, path of an alternative build of the same library (never a fallback)
LIB_PATH = os.environ.get("CPB200_LIB") or os.path.join(_HERE, "libcpb200.so")

# status codes / flags (include/cpb200.h)
CPB_OK = 0
CPB_ERR_INVALID = -1
CPB_ERR_CUDA = -2
CPB_ERR_NOMEM = -3
CPB_ERR_UNSUPPORTED = -4
CPB_ERR_CHARGE = -5
CPB_VPSI_OVERWRITE = 1
CPB_VPSI_TKSHAM = 2
CPB_RHO_CHECK_CHARGE = 1
CPB_RHO_ACCUMULATE = 2
CPB_C0_KEEP = 0x10
CPB_C0_REUSE = 0x20
CPB_PSI_KEEP = 0x40
CPB_PSI_REUSE = 0x80
CPB_ASYNC = 0x100
CPB_DENSE_ACCUMULATE = 1
CPB_PEER_HANDLE_BYTES = 64


class PlanInfo(C.Structure):
    _fields_ = [
        ("nr", C.c_int * 3),
        ("kr", C.c_int * 3),
        ("ngw", C.c_int),
        ("geq0", C.c_int),
        ("nrays", C.c_int),
        ("zband", C.c_int),
        ("xband", C.c_int),
        ("max_batch", C.c_int),
        ("device", C.c_int),
        ("radix", (C.c_int * 2) * 3),
        ("workspace_bytes", C.c_size_t),
        ("band_pruned", C.c_int * 3),
        ("chunk_xtiles", C.c_int),
        ("streams", C.c_int),
        ("z_warp_kernels", C.c_int),
        ("z_warp_radix", C.c_int * 2),
        ("x_warp_kernels", C.c_int),
        ("x_warp_radix", C.c_int),
    ]


#: every symbol include/cpb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cpb_last_error": (C.c_char_p, []),
    "cpb_version": (C.c_char_p, []),
    "cpb_length_supported": (C.c_int, [C.c_int]),
    "cpb_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]),
    "cpb_plan_destroy": (C.c_int, [C.c_void_p]),
    "cpb_plan_get_info": (C.c_int, [C.c_void_p, C.POINTER(PlanInfo)]),
    "cpb_plan_get_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cpb_part_1d_nbr_el_in_blk": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "cpb_part_1d_get_el_in_blk": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cpb_rhoofr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_int,
                             C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                             C.POINTER(C.c_double), C.c_uint]),
    "cpb_vpsi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_int, C.c_int, C.c_uint]),
    "cpb_rhoofr_lsd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint]),
    "cpb_vpsi_lsd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_int, C.c_uint]),
    "cpb_rhoofr_lsd_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.c_uint, C.c_void_p]),
    "cpb_vpsi_lsd_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "cpb_lsd_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_void_p]),
    "cpb_c0_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int]),
    "cpb_c0_invalidate": (C.c_int, [C.c_void_p]),
    "cpb_rhoofr_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_uint, C.c_void_p]),
    "cpb_vpsi_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "cpb_dense_fwfft_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p]),
    "cpb_dense_invfft_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_uint, C.c_void_p]),
    "cpb_vofrho_local_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "cpb_vofrho_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "cpb_rhoofr_kpt_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_double, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint, C.c_void_p]),
    "cpb_vpsi_kpt_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "cpb_peer_last_error": (C.c_char_p, []),
    "cpb_peer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p]),
    "cpb_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cpb_peer_local_ptr": (C.c_void_p, [C.c_void_p]),
    "cpb_peer_barrier": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cpb_peer_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cpb_peer_allreduce_f64": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "cpb_peer_bcast_f64": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]),
    "cpb_peer_allgather_f64": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]),
    "cpb_peer_redist_c2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_long, C.c_int, C.c_void_p]),
    "cpb_peer_allreduce_scalars": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_void_p]),
    "cpb_peer_set_timeout_ms": (C.c_int, [C.c_void_p, C.c_double]),
    "cpb_peer_destroy": (C.c_int, [C.c_void_p]),
    "cpb_rhoofr_kpt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_double, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint]),
    "cpb_vpsi_kpt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint]),
    "cpb_tauofr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                             C.c_int, C.c_void_p, C.c_uint]),
    "cpb_vtaupsi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint]),
    "cpb_tauofr_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p, C.c_uint, C.c_void_p]),
    "cpb_vtaupsi_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "cpb_hfx_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint, C.c_void_p]),
    "cpb_hfx": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                          C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint]),
    "cpb_plan_launch_count": (C.c_long, [C.c_void_p]),
    "cpb_plan_set_vpot_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cpb_rhoofr_finish": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_double)] * 5),
    "cpb_rhoofr_pending": (C.c_int, [C.c_void_p]),
    "cpb_plan_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "cpb_plan_set_streams": (C.c_int, [C.c_void_p, C.c_int]),
    "cpb_plan_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]),
}

KERNEL_KINDS = ("x_inv", "y_inv", "z_rho", "z_vpsi", "y_fwd", "x_fwd", "kin_energy", "rho_sum", "unpack",
                "dense")


def declare(cdll: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every exported entry point; raises AttributeError if one is
    missing from the shared object."""
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


class LibraryNotBuilt(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library.  No fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryNotBuilt(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cpmd_b200/csrc`. cpmd_b200 has no CPU fallback.")
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib
