"""ctypes loader of oracle/_ref/libref_gpu.so: the REFERENCE's own GPU path for rhoofr / vpsi (its CUDA
sources src/cuuser_utils.cu + src/cuuser_utils_kernels.cu compiled by nvcc where they lie, cuFFT plans and
stage order per cp_cufft_utils / fftcu_methods - see oracle/ref_gpu_driver.cu).

TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ as a device-side parity check made of reference code and
by bench.py's `reference_gpu` leg.  ``load()`` returns None when the library has not been built (it needs
/root/reference at build time; the GPU box uses the prebuilt file that travels with the snapshot)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_gpu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH) and os.path.exists("/root/reference/src/cuuser_utils.cu"):
            import subprocess
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.call(["make", "-C", _HERE, "_ref/libref_gpu.so"], env=env, stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL)
        if not os.path.exists(PATH):
            return None
        try:
            L = C.CDLL(PATH)
        except OSError:
            return None
        vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
        L.refgpu_create.argtypes = [C.POINTER(vp), vp, vp, i, i, vp, vp, vp, i, i, vp, i, d, d]
        L.refgpu_destroy.argtypes = [vp]
        L.refgpu_rhoofr.argtypes = [vp, vp, l, i, vp, vp, i]
        L.refgpu_vpsi.argtypes = [vp, vp, vp, l, i, vp, vp, i]
        L.refgpu_launches.argtypes = [vp]
        L.refgpu_launches.restype = l
        L.refgpu_last_error.restype = C.c_char_p
        L.refgpu_source.restype = C.c_char_p
        _lib = L
    return _lib


class RefGpu:
    """One task, one device, one stream of the reference's GPU path on geometry ``geo`` (oracle Geometry)."""

    def __init__(self, geo, tpiba2=1.0, omega=1.0):
        self.L = load()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libref_gpu.so not built")
        self.geo = geo
        self.h = C.c_void_p()
        nr = np.asarray(geo.nr, dtype=np.int32)
        kr = np.asarray(geo.kr, dtype=np.int32)
        self._keep = [np.ascontiguousarray(a, dtype=np.int32) for a in (geo.nzhs, geo.indzs, geo.msp2)]
        hg = np.ascontiguousarray(geo.hg, dtype=np.float64)
        rc = self.L.refgpu_create(C.byref(self.h), nr.ctypes.data, kr.ctypes.data, geo.ngw, geo.nrays,
                                  self._keep[0].ctypes.data, self._keep[1].ctypes.data, self._keep[2].ctypes.data,
                                  geo.kr3min, geo.kr3max, hg.ctypes.data, int(geo.geq0), float(tpiba2), float(omega))
        if rc:
            raise RuntimeError(self.L.refgpu_last_error().decode())

    def rhoofr(self, c0, f, device_scatter=False):
        c0 = np.ascontiguousarray(c0, dtype=np.complex128)
        f = np.ascontiguousarray(f, dtype=np.float64)
        rho = np.empty(self.geo.nnr1)
        if self.L.refgpu_rhoofr(self.h, c0.ctypes.data, c0.shape[1], c0.shape[0], f.ctypes.data, rho.ctypes.data,
                                int(device_scatter)):
            raise RuntimeError(self.L.refgpu_last_error().decode())
        return rho

    def vpsi(self, c0, c2, f, vpot, device_scatter=False):
        c0 = np.ascontiguousarray(c0, dtype=np.complex128)
        out = np.array(c2, dtype=np.complex128, order="C", copy=True)
        f = np.ascontiguousarray(f, dtype=np.float64)
        v = np.ascontiguousarray(vpot, dtype=np.float64)
        if self.L.refgpu_vpsi(self.h, c0.ctypes.data, out.ctypes.data, c0.shape[1], c0.shape[0], f.ctypes.data,
                              v.ctypes.data, int(device_scatter)):
            raise RuntimeError(self.L.refgpu_last_error().decode())
        return out

    @property
    def launches(self):
        return int(self.L.refgpu_launches(self.h))

    def close(self):
        if self.h:
            self.L.refgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
