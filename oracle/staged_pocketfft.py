"""Staged sparse CPU restatement of vpsi + rhoofr with a LIBRARY 1-D FFT (scipy.fft = pocketfft,
threaded) in the role FFTW plays in the reference's "FFTW build" (mltfft_fftw,
mltfft_utils.mod.F90:492-609).

TEST INFRASTRUCTURE / CPU baseline only (see oracle/cpmd_oracle.py header): used by tests/ as a third
restatement and by bench.py's cpu_baseline / --impl reference legs as the stronger timed CPU path.
PARITY UNPINNED by reference fixtures (none exist); checked against the dense NumPy restatement.

The passes are fftnew's sparse branches step by step (paths relative to /root/reference/src):
  set_psi_2_states_g / set_psi_1_state_g                     state_utils.mod.F90:132-189
  invfftn: x mltfft over the msrays rays -> unpack_x2y (zero + scatter through msp(:,2)) -> y mltfft
           over the z band -> putz -> z mltfft                fftmain_utils.mod.F90:92-104
  fwfftn:  the mirror, scale 1/(n1 n2 n3)                     fftmain_utils.mod.F90:122-136
  build_density_sum                                           density_utils.mod.F90:61-83
  V psi, unpack + kinetic + occupation scale                  vpsi_utils.mod.F90:487-493, 626-673
  kin_energy / dotp                                           kin_energy_utils.mod.F90:62-110
Several pairs are transformed per library call (one batched call per pass) so that the Python
overhead is negligible next to the transforms.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sf

from . import cpmd_oracle as orc

_WORKERS = -1


def set_threads(n=0):
    """n <= 0: every core this process may run on (ignores OMP_NUM_THREADS, which torchrun sets to 1).
    Sets both the pocketfft worker count and the OpenMP team of the copy helpers."""
    global _WORKERS
    from . import staged
    _WORKERS = int(staged.set_threads(n))
    return _WORKERS


_H = None


def _helpers():
    """oracle/staged_helpers.c (built into libstaged_oracle.so): the threaded copy / scatter / pointwise
    steps between the library FFT calls."""
    global _H
    if _H is None:
        import ctypes as C

        from . import staged
        L = staged.load()
        vp, i, l, d = C.c_void_p, C.c_int, C.c_long, C.c_double
        for name, args in (("orc_h_set_psi", [i, l, l, vp, vp, vp, vp, vp]),
                           ("orc_h_unpack_x2y", [i, l, l, l, vp, vp, vp]),
                           ("orc_h_pack_y2x", [i, l, l, l, vp, vp, vp]),
                           ("orc_h_putz", [i, l, l, l, l, vp, vp]),
                           ("orc_h_density_sum", [i, l, vp, vp, vp, vp]),
                           ("orc_h_vmul", [i, l, vp, vp]),
                           ("orc_h_gather_g", [i, l, l, vp, vp, vp, d, vp, vp])):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = args
        _H = L
    return _H


_MAPS = {}


def _maps(geo):
    key = id(geo)
    hit = _MAPS.get(key)
    m = hit[1] if hit is not None and hit[0] is geo else None   # id() alone is reused after garbage collection
    if m is None:
        n1, n2, n3 = geo.nr
        kr1, kr2 = geo.kr[0], geo.kr[1]
        nzb = geo.kr3max - geo.kr3min + 1
        nz = geo.nzhs.astype(np.int64) - 1          # into (nrays, kr1s) ray storage, x fastest
        iz = geo.indzs.astype(np.int64) - 1
        nzc = np.ascontiguousarray((nz // kr1) * n1 + nz % kr1)   # same position in unpadded (nrays, n1)
        izc = np.ascontiguousarray((iz // kr1) * n1 + iz % kr1)
        ms = np.ascontiguousarray(geo.msp2.astype(np.int64) - 1)  # ray -> y + (z - kr3min) * kr2s
        m = (n1, n2, n3, kr2, nzb, nzc, izc, ms)
        _MAPS.clear()
        _MAPS[key] = (geo, m)      # holds the geometry: its id cannot be handed to another object meanwhile
    return m


def _inv_batch(geo, a, b):
    """a, b: (np, ngw) coefficients of the two states of each pair (b zero for a single state).
    Returns psi (np, n3, n2, n1) = invfftn of the packed pairs (fftmain_utils.mod.F90:92-104)."""
    H = _helpers()
    n1, n2, n3, kr2, nzb, nzc, izc, ms = _maps(geo)
    npair = a.shape[0]
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    rays = np.empty((npair, geo.nrays, n1), dtype=np.complex128)
    H.orc_h_set_psi(npair, geo.ngw, geo.nrays * n1, a.ctypes.data, b.ctypes.data, nzc.ctypes.data, izc.ctypes.data,
                    rays.ctypes.data)                                     # zeroing + state_utils.mod.F90:171-189
    x = sf.ifft(rays, axis=2, norm="forward", workers=_WORKERS, overwrite_x=True)
    yf = np.empty((npair, nzb, kr2, n1), dtype=np.complex128)
    H.orc_h_unpack_x2y(npair, geo.nrays, nzb * kr2, n1, ms.ctypes.data, x.ctypes.data, yf.ctypes.data)
    y = sf.ifft(yf[:, :, :n2, :], axis=2, norm="forward", workers=_WORKERS)
    y = np.ascontiguousarray(y)
    full = np.empty((npair, n3, n2, n1), dtype=np.complex128)
    H.orc_h_putz(npair, n3, geo.kr3min - 1, nzb, n2 * n1, y.ctypes.data, full.ctypes.data)
    return sf.ifft(full, axis=1, norm="forward", workers=_WORKERS, overwrite_x=True)


def _fwd_batch(geo, psi):
    """fwfftn of (np, n3, n2, n1) arrays (fftmain_utils.mod.F90:122-136); returns psi(+G), psi(-G) (np, ngw)."""
    H = _helpers()
    n1, n2, n3, kr2, nzb, nzc, izc, ms = _maps(geo)
    npair = psi.shape[0]
    z = sf.fft(psi, axis=1, workers=_WORKERS, overwrite_x=True)[:, geo.kr3min - 1:geo.kr3max]   # getz
    yf = np.empty((npair, nzb, kr2, n1), dtype=np.complex128)
    yf[:, :, :n2, :] = sf.fft(z, axis=2, workers=_WORKERS)
    x = np.empty((npair, geo.nrays, n1), dtype=np.complex128)
    H.orc_h_pack_y2x(npair, geo.nrays, nzb * kr2, n1, ms.ctypes.data, yf.ctypes.data, x.ctypes.data)
    rays = np.ascontiguousarray(sf.fft(x, axis=2, workers=_WORKERS, overwrite_x=True))
    pp = np.empty((npair, geo.ngw), dtype=np.complex128)
    pm = np.empty((npair, geo.ngw), dtype=np.complex128)
    H.orc_h_gather_g(npair, geo.ngw, geo.nrays * n1, nzc.ctypes.data, izc.ctypes.data, rays.ctypes.data,
                     1.0 / (float(n1) * n2 * n3), pp.ctypes.data, pm.ctypes.data)
    return pp, pm


def _pairs(nstate, group, ngroups):
    return orc.state_pairs(nstate, group, ngroups)


def rhoofr(geo, c0, f, omega, tpiba2, group=0, ngroups=1, batch=8):
    """Same contract as cpmd_oracle.rhoofr (rhoofr_utils.mod.F90:122-644, Gamma point, no LSD)."""
    n1, n2, n3 = geo.nr
    H = _helpers()
    dense = np.zeros((n3, n2, n1))
    pairs = [(i, j) for (i, j) in _pairs(c0.shape[0], group, ngroups)
             if f[i] != 0.0 or (j is not None and f[j] != 0.0)]           # :312-316
    for o in range(0, len(pairs), batch):
        pb = pairs[o:o + batch]
        a = np.stack([c0[i, :geo.ngw] for i, _ in pb])
        b = np.stack([c0[j, :geo.ngw] if j is not None else np.zeros(geo.ngw, complex) for _, j in pb])
        psi = _inv_batch(geo, a, b)
        ca = np.array([f[i] / omega for i, _ in pb])
        cb = np.array([f[j] / omega if j is not None else 0.0 for _, j in pb])
        psi = np.ascontiguousarray(psi)
        H.orc_h_density_sum(len(pb), n1 * n2 * n3, psi.ctypes.data, ca.ctypes.data, cb.ctypes.data,
                            dense.ctypes.data)                            # density_utils.mod.F90:61-83
    rho = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]))
    rho[:n3, :n2, :n1] = dense
    blk = range(c0.shape[0])           # the reference sums over ALL states on every group (:178)
    w = np.full(geo.ngw, 2.0)
    if geo.geq0:
        w[0] = 1.0
    ekin = rsum = 0.0
    for i in blk:                                                         # kin_energy_utils.mod.F90:62-110
        if f[i] != 0.0:
            c = c0[i, :geo.ngw]
            m = c.real ** 2 + c.imag ** 2
            if geo.geq0:
                m0 = m.copy()
                m0[0] = c[0].real ** 2                                    # dotp_utils.mod.F90:26-53
            else:
                m0 = m
            rsum += f[i] * float(np.dot(w, m0))
            ekin += f[i] * float(np.dot(geo.hg, m)) * tpiba2
    rsum_r = rho.sum() * omega / (float(n1) * n2 * n3)
    return dict(rhoe=rho.reshape(-1), ekin=ekin, rsum_g=rsum, rsum_r=rsum_r)


def vpsi(geo, c0, c2, f, vpot, tpiba2, group=0, ngroups=1, tksham=False, batch=8):
    """Same contract as cpmd_oracle.vpsi (vpsi_utils.mod.F90:120-732, Gamma point, RKS); returns the new c2."""
    n1, n2, n3 = geo.nr
    out = np.array(c2, dtype=np.complex128, copy=True)
    H = _helpers()
    v = np.ascontiguousarray(np.asarray(vpot).reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n3, :n2, :n1])
    g2 = tpiba2 * geo.hg
    pairs = _pairs(c0.shape[0], group, ngroups)
    for o in range(0, len(pairs), batch):
        pb = pairs[o:o + batch]
        a = np.stack([c0[i, :geo.ngw] for i, _ in pb])
        b = np.stack([c0[j, :geo.ngw] if j is not None else np.zeros(geo.ngw, complex) for _, j in pb])
        psi = _inv_batch(geo, a, b)
        psi = np.ascontiguousarray(psi)
        H.orc_h_vmul(len(pb), n1 * n2 * n3, psi.ctypes.data, v.ctypes.data)   # :487-493
        pp, pm = _fwd_batch(geo, psi)
        fp, fm = pp + pm, pp - pm                                         # :655-671
        for q, (i, j) in enumerate(pb):
            fi = 0.5 * f[i]
            if fi == 0.0:
                fi = 0.5 if tksham else 1.0                               # :627-633
            out[i, :geo.ngw] += -fi * ((g2 * a[q].real + fp[q].real) + 1j * (g2 * a[q].imag + fm[q].imag))
            if j is not None:
                fj = 0.5 * f[j]
                if fj == 0.0:
                    fj = 0.5 if tksham else 1.0
                out[j, :geo.ngw] += -fj * ((g2 * b[q].real + fp[q].imag) + 1j * (g2 * b[q].imag - fm[q].real))
    return out
