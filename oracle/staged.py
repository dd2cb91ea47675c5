"""ctypes wrapper of oracle/staged_oracle.c (TEST INFRASTRUCTURE / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libstaged_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            env = dict(os.environ)
            env.pop("CC", None)
            subprocess.check_call(["make", "-C", _HERE], env=env, stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        L.orc_set_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        vp = C.c_void_p
        L.orc_rhoofr.restype = C.c_int
        L.orc_rhoofr.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int,
                                 C.c_double, C.c_double, vp, C.c_long, C.c_int, vp, C.c_int, C.c_int, vp,
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_vpsi.restype = C.c_int
        L.orc_vpsi.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int,
                               C.c_double, vp, vp, C.c_long, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def host_cores():
    """Cores this process may run on (sched affinity), NOT OMP_NUM_THREADS: torchrun exports
    OMP_NUM_THREADS=1, which made the round-1 reference arm run on one core at N > 1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n=0):
    """n <= 0: one thread per core of the affinity mask."""
    if n <= 0:
        n = host_cores()
    return load().orc_set_threads(int(n))


def _geo_args(geo):
    nr = np.asarray(geo.nr, dtype=np.int32)
    kr = np.asarray(geo.kr, dtype=np.int32)
    nz = np.ascontiguousarray(geo.nzhs, dtype=np.int32)
    iz = np.ascontiguousarray(geo.indzs, dtype=np.int32)
    ms = np.ascontiguousarray(geo.msp2, dtype=np.int32)
    hg = np.ascontiguousarray(geo.hg, dtype=np.float64)
    return nr, kr, nz, iz, ms, hg


def rhoofr(geo, c0, f, omega, tpiba2, group=0, ngroups=1):
    """Same contract as cpmd_oracle.rhoofr; c0 is (nstate, ld) C-contiguous."""
    L = load()
    nr, kr, nz, iz, ms, hg = _geo_args(geo)
    c0 = np.ascontiguousarray(c0, dtype=np.complex128)
    f = np.ascontiguousarray(f, dtype=np.float64)
    rhoe = np.empty(geo.nnr1, dtype=np.float64)
    ekin, rg, rr = C.c_double(), C.c_double(), C.c_double()
    rc = L.orc_rhoofr(nr.ctypes.data, kr.ctypes.data, geo.ngw, geo.nrays, nz.ctypes.data, iz.ctypes.data,
                      ms.ctypes.data, geo.kr3min, geo.kr3max, hg.ctypes.data, int(geo.geq0), tpiba2, omega,
                      c0.ctypes.data, c0.shape[1], c0.shape[0], f.ctypes.data, ngroups, group,
                      rhoe.ctypes.data, C.byref(ekin), C.byref(rg), C.byref(rr))
    if rc:
        raise RuntimeError(f"orc_rhoofr failed: {rc}")
    return dict(rhoe=rhoe, ekin=ekin.value, rsum_g=rg.value, rsum_r=rr.value)


def vpsi(geo, c0, c2, f, vpot, tpiba2, group=0, ngroups=1, tksham=False):
    L = load()
    nr, kr, nz, iz, ms, hg = _geo_args(geo)
    c0 = np.ascontiguousarray(c0, dtype=np.complex128)
    out = np.array(c2, dtype=np.complex128, order="C", copy=True)
    f = np.ascontiguousarray(f, dtype=np.float64)
    v = np.ascontiguousarray(vpot, dtype=np.float64)
    rc = L.orc_vpsi(nr.ctypes.data, kr.ctypes.data, geo.ngw, geo.nrays, nz.ctypes.data, iz.ctypes.data,
                    ms.ctypes.data, geo.kr3min, geo.kr3max, hg.ctypes.data, int(geo.geq0), tpiba2,
                    c0.ctypes.data, out.ctypes.data, c0.shape[1], c0.shape[0], f.ctypes.data, v.ctypes.data,
                    ngroups, group, int(tksham))
    if rc:
        raise RuntimeError(f"orc_vpsi failed: {rc}")
    return out
