"""ctypes loader of oracle/_ref/libcuuser_ref.so: the REFERENCE's own helper kernels
(/root/reference/src/cuuser_utils_kernels.cu) compiled for the host by oracle/Makefile.

TEST INFRASTRUCTURE ONLY (see the package docstring of oracle/cpmd_oracle.py).  ``load()`` returns
None when the library has not been built (no /root/reference and no prebuilt file): tests skip."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libcuuser_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH) and os.path.exists("/root/reference/src/cuuser_utils_kernels.cu"):
            # authoring container: build on demand (the GPU box has no /root/reference and uses the
            # prebuilt file that travels with the snapshot)
            import subprocess
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.call(["make", "-C", _HERE, "_ref/libcuuser_ref.so"], env=env, stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL)
        if not os.path.exists(PATH):
            return None
        L = C.CDLL(PATH)
        vp, ip, dbl, i = C.c_void_p, C.c_void_p, C.c_double, C.c_int
        L.ref_set_psi_2_states_g.argtypes = [vp, vp, vp, i, ip, ip, i]
        L.ref_set_psi_1_state_g.argtypes = [dbl, dbl, vp, vp, i, ip, ip, i]
        L.ref_build_density_sum.argtypes = [dbl, dbl, vp, vp, i]
        L.ref_pointwise_cxr.argtypes = [vp, vp, i]
        L.ref_phasen.argtypes = [vp, i, i, i, i, i, i, i]
        L.ref_putz.argtypes = [vp, vp, i, i, i, i]
        L.ref_getz.argtypes = [vp, vp, i, i, i, i]
        L.ref_unpack_x2y.argtypes = [vp, vp, i, i, i, ip, i, ip, i, i]
        L.ref_pack_y2x.argtypes = [vp, vp, i, i, i, ip, i, ip, i, i]
        L.ref_setblock2zero.argtypes = [vp, i, i, i, i, i]
        L.ref_source.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data


def set_psi_2_states_g(geo, c1, c2):
    """CuUser_Kernel_Set_Psi_2_Stages_G on zeroed ray storage psi(kr1s*nrays)."""
    L = load()
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    c1 = np.ascontiguousarray(c1, dtype=np.complex128)
    c2 = np.ascontiguousarray(c2, dtype=np.complex128)
    nz = np.ascontiguousarray(geo.nzhs, dtype=np.int32)
    iz = np.ascontiguousarray(geo.indzs, dtype=np.int32)
    L.ref_set_psi_2_states_g(_p(c1), _p(c2), _p(psi), geo.ngw, _p(nz), _p(iz), int(geo.geq0))
    return psi


def set_psi_1_state_g(geo, c1, alpha=1.0 + 0.0j):
    L = load()
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    c1 = np.ascontiguousarray(c1, dtype=np.complex128)
    nz = np.ascontiguousarray(geo.nzhs, dtype=np.int32)
    iz = np.ascontiguousarray(geo.indzs, dtype=np.int32)
    a = complex(alpha)
    L.ref_set_psi_1_state_g(a.real, a.imag, _p(c1), _p(psi), geo.ngw, _p(nz), _p(iz), int(geo.geq0))
    return psi


def build_density_sum(alpha_re, alpha_im, psi, rho):
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    load().ref_build_density_sum(float(alpha_re), float(alpha_im), _p(psi), _p(rho), psi.size)
    return rho


def pointwise_cxr(psi, v):
    out = np.array(psi, dtype=np.complex128)
    v = np.ascontiguousarray(v, dtype=np.float64)
    load().ref_pointwise_cxr(_p(out), _p(v), out.size)
    return out


def phasen(geo, f):
    """CuUser_Kernel_PhaseN with the arguments fftnew passes for one task: (kr1, kr2s, kr3s, n1u = 1,
    n1o = nr1, nr2s, nr3s) (fftmain_utils.mod.F90:117-119)."""
    out = np.array(f, dtype=np.complex128)
    n1, n2, n3 = geo.nr
    load().ref_phasen(_p(out), geo.kr[0], geo.kr[1], geo.kr[2], 1, n1, n2, n3)
    return out


def putz(a, krmin, krmax, kr, m):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.full(kr * m, 7.0 + 7.0j)
    load().ref_putz(_p(a), _p(b), krmin, krmax, kr, m)
    return b


def getz(a, krmin, krmax, kr, m):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.zeros((krmax - krmin + 1) * m, dtype=np.complex128)
    load().ref_getz(_p(a), _p(b), krmin, krmax, kr, m)
    return b


def setblock2zero(b, trans, n, m, ldbx, ldby):
    """CuUser_C_SetBlock2Zero(b, transb, n, m, ldbx, ldby): what mltfft_cuda does to its output
    (mltfft_utils.mod.F90:643-646).  In place on a complex128 array of ldbx*ldby elements."""
    assert b.dtype == np.complex128 and b.size == ldbx * ldby
    load().ref_setblock2zero(_p(b), 1 if trans in ("N", "n") else 0, n, m, ldbx, ldby)
    return b


def unpack_x2y(geo, xf, lr1):
    """CuUser_Kernel_Unpack_x2y_8 for one task: xf[ray + nrays*x] -> yf[x*mm + msp(ray) - 1], mm =
    kr2s*(kr3max-kr3min+1); yf zeroed first as unpack_x2y does (fftutil_utils.mod.F90:413)."""
    mm = geo.kr[1] * (geo.kr3max - geo.kr3min + 1)
    xf = np.ascontiguousarray(xf, dtype=np.complex128)
    yf = np.zeros(mm * lr1, dtype=np.complex128)
    msp = np.ascontiguousarray(geo.msp2, dtype=np.int32)
    sp8 = np.array([geo.nrays], dtype=np.int32)
    load().ref_unpack_x2y(_p(xf), _p(yf), mm, lr1, geo.nrays * lr1, _p(msp), geo.nrays, _p(sp8), 0, 1)
    return yf


def pack_y2x(geo, yf, lr1):
    mm = geo.kr[1] * (geo.kr3max - geo.kr3min + 1)
    yf = np.ascontiguousarray(yf, dtype=np.complex128)
    xf = np.zeros(geo.nrays * lr1, dtype=np.complex128)
    msp = np.ascontiguousarray(geo.msp2, dtype=np.int32)
    sp8 = np.array([geo.nrays], dtype=np.int32)
    load().ref_pack_y2x(_p(xf), _p(yf), mm, lr1, geo.nrays * lr1, _p(msp), geo.nrays, _p(sp8), 0, 1)
    return xf
