"""Host-side mirror of the reference interfaces for the vpsi / rhoofr path.

Two levels:

* :class:`Plan` — thin object over the C ABI (``include/cpb200.h``): host-array entry points
  (what the Fortran shim binds) and device-pointer entry points (torch CUDA tensors).
* :class:`CpmdContext` — mirrors the *reference's own call signatures*: the reference keeps mesh,
  G-vector and occupation data in module globals (``spar``, ``fpar``, ``ncpw``, ``cppt``, ``parm``,
  ``crge``, ``parai``) and calls ``rhoofr(c0,rhoe,psi,nstate)`` (rhoofr_utils.mod.F90:122) and
  ``vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2)`` (vpsi_utils.mod.F90:120); the context
  holds those globals and exposes methods with exactly those argument lists.  Unsupported
  variants raise :class:`StopGM` like the reference's ``stopgm``.

Array conventions: a Fortran array ``c0(ld, nstate)`` is a C-contiguous numpy/torch array of
shape ``(nstate, ld)``; ``rhoe``/``vpot`` ``(nnr1[, nspin])`` are flat float64 arrays of length
``kr1*kr2s*kr3s`` (x fastest).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib


class CpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cpb200 error {code}: {msg}")
        self.code = code


class StopGM(RuntimeError):
    """Python stand-in for ``CALL stopgm(procedure, message, __LINE__, __FILE__)``
    (error_handling.mod.F90:11-53): the reference's only error convention is to abort."""

    def __init__(self, procedure, message):
        super().__init__(f"{procedure}: {message}")
        self.procedure = procedure
        self.message = message


def leadim(nr: int) -> int:
    """kr = nr + MOD(nr+1, 2) (loadpa_utils.mod.F90:509-525)."""
    return nr + (nr + 1) % 2


def _ptr(a):
    """Raw address of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch


def _is_torch(a):
    return not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


def _numel(a):
    return a.numel() if _is_torch(a) else np.asarray(a).size


class Plan:
    """One FFT/G-vector plan on one GPU (``cpb_plan_create``)."""

    def __init__(self, nr, inyh, hg, tpiba2=1.0, omega=1.0, kr=None, device=0, max_batch=32, _cdll=None):
        self._L = _cdll if _cdll is not None else _lib.load()
        self._h = C.c_void_p()
        nr = tuple(int(v) for v in nr)
        kr = tuple(leadim(v) for v in nr) if kr is None else tuple(int(v) for v in kr)
        inyh = np.asarray(inyh)
        if inyh.ndim != 2 or inyh.shape[0] != 3:
            raise ValueError("inyh must have shape (3, ngw) like the Fortran array")
        ngw = inyh.shape[1]
        inyh_f = np.ascontiguousarray(inyh.T, dtype=np.int32)  # (ngw,3) C-order == (3,ngw) Fortran
        hg = np.ascontiguousarray(hg, dtype=np.float64)
        if hg.shape != (ngw,):
            raise ValueError("hg must have shape (ngw,)")
        nr_c = (C.c_int * 3)(*nr)
        kr_c = (C.c_int * 3)(*kr)
        rc = self._L.cpb_plan_create(C.byref(self._h), nr_c, kr_c, ngw, inyh_f.ctypes.data, hg.ctypes.data,
                                     float(tpiba2), float(omega), int(device), int(max_batch))
        self._check(rc)
        self.nr, self.kr, self.ngw = nr, kr, ngw
        self.tpiba2, self.omega, self.device = float(tpiba2), float(omega), int(device)
        self.nnr1 = kr[0] * kr[1] * kr[2]

    # -- plumbing -------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise CpbError(rc, self._L.cpb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.cpb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self):
        inf = _lib.PlanInfo()
        self._check(self._L.cpb_plan_get_info(self._h, C.byref(inf)))
        return dict(nr=tuple(inf.nr), kr=tuple(inf.kr), ngw=inf.ngw, geq0=bool(inf.geq0), nrays=inf.nrays,
                    zband=inf.zband, xband=inf.xband, max_batch=inf.max_batch, device=inf.device,
                    radix=tuple((inf.radix[d][0], inf.radix[d][1]) for d in range(3)),
                    workspace_bytes=inf.workspace_bytes,
                    band_pruned=tuple(inf.band_pruned), chunk_xtiles=inf.chunk_xtiles, streams=inf.streams,
                    z_warp_kernels=bool(inf.z_warp_kernels), z_warp_radix=tuple(inf.z_warp_radix),
                    x_warp_kernels=inf.x_warp_kernels, x_warp_radix=inf.x_warp_radix)

    def maps(self):
        """(nzhs, indzs) with the reference's numbering (fftprp_utils.mod.F90:269-285)."""
        nzhs = np.empty(self.ngw, dtype=np.int32)
        indzs = np.empty(self.ngw, dtype=np.int32)
        self._check(self._L.cpb_plan_get_maps(self._h, nzhs.ctypes.data, indzs.ctypes.data))
        return nzhs, indzs

    @property
    def launch_count(self):
        return int(self._L.cpb_plan_launch_count(self._h))

    def set_vpot_event(self, event):
        """One-shot: the next ``vpsi*_dev`` call waits for ``event`` (torch.cuda.Event recorded on the
        stream that produces vpot, or a raw cudaEvent_t) only before its first z pass."""
        h = None if event is None else (event if isinstance(event, int) else event.cuda_event)
        self._check(self._L.cpb_plan_set_vpot_event(self._h, C.c_void_p(h) if h is not None else None))

    def set_profiling(self, on=True):
        self._check(self._L.cpb_plan_set_profiling(self._h, int(bool(on))))

    def set_streams(self, n):
        """Number of work spaces/streams the batches of a call alternate between (1 = serialised)."""
        self._check(self._L.cpb_plan_set_streams(self._h, int(n)))

    def kernel_times(self, reset=False):
        """{kernel class: (total ms, launches)} accumulated while profiling was on."""
        n = len(_lib.KERNEL_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_long * n)()
        self._check(self._L.cpb_plan_get_kernel_times(self._h, ms, cnt, int(bool(reset))))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_lib.KERNEL_KINDS)}

    def _c0_args(self, c0, nstate):
        if c0.ndim != 2:
            raise ValueError("c0 must be (nstate, ld)")
        ns, ld = c0.shape
        if nstate is None:
            nstate = ns
        if nstate > ns or ld < self.ngw:
            raise ValueError("c0 shape inconsistent with nstate/ngw")
        if _is_torch(c0) and not c0.is_contiguous():
            raise ValueError("c0 must be contiguous")
        return int(nstate), int(ld)

    @staticmethod
    def _f_arg(f, nstate):
        """occupations as a contiguous float64 host array with at least nstate entries (the C side
        indexes f[0 .. nstate))"""
        f = np.ascontiguousarray(f, dtype=np.float64)
        if f.ndim != 1 or f.shape[0] < nstate:
            raise ValueError(f"f must hold at least nstate = {nstate} occupations")
        return f

    def _g_arg(self, a, name, per=1):
        """a per-plane-wave real array (hgkp, hgkm: ngw; gk: 3*ngw) - host or device; None is passed
        on as NULL (the C side rejects it)"""
        if a is None:
            return a
        n = a.numel() if _is_torch(a) else np.asarray(a).size
        if n < per * self.ngw:
            raise ValueError(f"{name} must hold at least {per}*ngw = {per * self.ngw} doubles")
        if _is_torch(a) and not a.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        return a

    # -- host-array entry points (the Fortran drop-in path) --------------------------------
    def rhoofr(self, c0, f, rhoe=None, nstate=None, ngroups=1, my_group=0, flags=0):
        """``cpb_rhoofr``: c0 (nstate, ld) complex128 numpy (host).  Returns
        (rhoe, ekin, rsum_g, rsum_r); rhoe is written into the given array if supplied."""
        c0 = _as_host(c0, np.complex128)
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if rhoe is None:
            rhoe = np.empty(self.nnr1, dtype=np.float64)
        rh = _as_host(rhoe, np.float64)
        if rh.size < self.nnr1:
            raise ValueError("rhoe too small")
        ekin, rg, rr = C.c_double(), C.c_double(), C.c_double()
        rc = self._L.cpb_rhoofr(self._h, c0.ctypes.data, ld, nstate, f.ctypes.data, ngroups, my_group,
                                rh.ctypes.data, C.byref(ekin), C.byref(rg), C.byref(rr), flags)
        self._check(rc)
        return rhoe, ekin.value, rg.value, rr.value

    def vpsi(self, c0, c2, f, vpot, nstate=None, ngroups=1, my_group=0, flags=0):
        """``cpb_vpsi``: c2 (host, in/out) is accumulated into unless CPB_VPSI_OVERWRITE."""
        c0 = _as_host(c0, np.complex128)
        c2h = _as_host(c2, np.complex128)
        if c2h.shape != c0.shape:
            raise ValueError("c2 must have the shape of c0")
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        v = _as_host(vpot, np.float64)
        if v.size < self.nnr1:
            raise ValueError("vpot too small")
        rc = self._L.cpb_vpsi(self._h, c0.ctypes.data, c2h.ctypes.data, ld, nstate, f.ctypes.data,
                              v.ctypes.data, ngroups, my_group, flags)
        self._check(rc)
        return c2

    # -- LSD (cntl%tlsd): rhoe / vpot are (2, nnr1) arrays = Fortran (nnr1, 2) ---------------------
    def rhoofr_lsd(self, c0, f, nsup, rhoe=None, nstate=None, ngroups=1, my_group=0, flags=0):
        """``cpb_rhoofr_lsd`` (host arrays).  Returns (rhoe, ekin, rsum_g, rsum_r, csums, csumsabs)."""
        c0 = _as_host(c0, np.complex128)
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if rhoe is None:
            rhoe = np.empty((2, self.nnr1), dtype=np.float64)
        rh = _as_host(rhoe, np.float64)
        if rh.size < 2 * self.nnr1:
            raise ValueError("rhoe too small (needs two columns)")
        out = [C.c_double() for _ in range(5)]
        rc = self._L.cpb_rhoofr_lsd(self._h, c0.ctypes.data, ld, nstate, f.ctypes.data, int(nsup), ngroups, my_group,
                                    rh.ctypes.data, *[C.byref(o) for o in out], flags)
        self._check(rc)
        return (rhoe, *[o.value for o in out])

    def vpsi_lsd(self, c0, c2, f, nsup, vpot, nstate=None, ngroups=1, my_group=0, flags=0):
        """``cpb_vpsi_lsd`` (host arrays); vpot is (2, nnr1): [alpha, beta]."""
        c0 = _as_host(c0, np.complex128)
        c2h = _as_host(c2, np.complex128)
        if c2h.shape != c0.shape:
            raise ValueError("c2 must have the shape of c0")
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        v = _as_host(vpot, np.float64)
        if v.size < 2 * self.nnr1:
            raise ValueError("vpot too small (needs two columns)")
        rc = self._L.cpb_vpsi_lsd(self._h, c0.ctypes.data, c2h.ctypes.data, ld, nstate, f.ctypes.data, int(nsup),
                                  v.ctypes.data, ngroups, my_group, flags)
        self._check(rc)
        return c2

    def rhoofr_lsd_dev(self, c0, f, nsup, rhoe, nstate=None, ngroups=1, my_group=0, flags=0, stream=None):
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if rhoe.numel() < 2 * self.nnr1:
            raise ValueError("rhoe too small (needs two columns)")
        out = [C.c_double() for _ in range(5)]
        rc = self._L.cpb_rhoofr_lsd_dev(self._h, _ptr(c0), ld, nstate, f.ctypes.data, int(nsup), ngroups, my_group,
                                        _ptr(rhoe), *[C.byref(o) for o in out], flags, _stream_ptr(stream))
        self._check(rc)
        return tuple(o.value for o in out)

    def vpsi_lsd_dev(self, c0, c2, f, nsup, vpot, nstate=None, ngroups=1, my_group=0, flags=0, stream=None):
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if vpot.numel() < 2 * self.nnr1:
            raise ValueError("vpot too small (needs two columns)")
        rc = self._L.cpb_vpsi_lsd_dev(self._h, _ptr(c0), _ptr(c2), ld, nstate, f.ctypes.data, int(nsup), _ptr(vpot),
                                      ngroups, my_group, flags, _stream_ptr(stream))
        self._check(rc)
        return c2

    def lsd_finish_dev(self, rhoe, stream=None):
        """rhoofr_utils.mod.F90:543-559 on group-summed channel densities; returns (rsum_r, csums, csumsabs)."""
        out = [C.c_double() for _ in range(3)]
        self._check(self._L.cpb_lsd_finish_dev(self._h, _ptr(rhoe), *[C.byref(o) for o in out], _stream_ptr(stream)))
        return tuple(o.value for o in out)

    def c0_upload(self, c0, nstate=None, ngroups=1, my_group=0):
        c0 = _as_host(c0, np.complex128)
        nstate, ld = self._c0_args(c0, nstate)
        self._check(self._L.cpb_c0_upload(self._h, c0.ctypes.data, ld, nstate, ngroups, my_group))

    def c0_invalidate(self):
        self._check(self._L.cpb_c0_invalidate(self._h))

    # -- device-pointer entry points (torch CUDA tensors) ----------------------------------
    def rhoofr_dev(self, c0, f, rhoe, nstate=None, ngroups=1, my_group=0, flags=0, stream=None):
        """``cpb_rhoofr_dev``: c0 (nstate, ld) complex128 CUDA tensor, rhoe float64 CUDA tensor of
        nnr1 elements (overwritten).  Returns (ekin, rsum_g, rsum_r) of the group's block; with
        ``lib.CPB_ASYNC`` in ``flags`` the call only enqueues and returns None: :meth:`rhoofr_finish`
        hands out the three sums."""
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if rhoe.numel() < self.nnr1:
            raise ValueError("rhoe too small")
        ekin, rg, rr = C.c_double(), C.c_double(), C.c_double()
        rc = self._L.cpb_rhoofr_dev(self._h, _ptr(c0), ld, nstate, f.ctypes.data, ngroups, my_group,
                                    _ptr(rhoe), C.byref(ekin), C.byref(rg), C.byref(rr), flags,
                                    _stream_ptr(stream))
        self._check(rc)
        if flags & _lib.CPB_ASYNC and self._L.cpb_rhoofr_pending(self._h):
            return None
        return ekin.value, rg.value, rr.value

    def rhoofr_finish(self):
        """``cpb_rhoofr_finish``: (ekin, rsum_g, rsum_r) of the pending ``CPB_ASYNC`` :meth:`rhoofr_dev` call;
        waits for that call's partial sums only."""
        ekin, rg, rr = C.c_double(), C.c_double(), C.c_double()
        self._check(self._L.cpb_rhoofr_finish(self._h, C.byref(ekin), C.byref(rg), C.byref(rr), None, None))
        return ekin.value, rg.value, rr.value

    def vpsi_dev(self, c0, c2, f, vpot, nstate=None, ngroups=1, my_group=0, flags=0, stream=None):
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if tuple(c2.shape) != tuple(c0.shape) or _numel(vpot) < self.nnr1:
            raise ValueError("c2 must have the shape of c0 and vpot nnr1 entries")
        rc = self._L.cpb_vpsi_dev(self._h, _ptr(c0), _ptr(c2), ld, nstate, f.ctypes.data, _ptr(vpot),
                                  ngroups, my_group, flags, _stream_ptr(stream))
        self._check(rc)
        return c2

    # -- k-points (tkpts%tkpnt): c0/c2 (nstate, ld >= 2 ngw) CUDA tensors, one k-point per call ------
    def _kpt_args(self, c0, nstate):
        if c0.ndim != 2:
            raise ValueError("c0 must be (nstate, ld)")
        ns, ld = c0.shape
        if nstate is None:
            nstate = ns
        if nstate > ns or ld < 2 * self.ngw:
            raise ValueError("k-point c0 must be (nstate, ld >= 2*ngw)")
        return int(nstate), int(ld)

    def rhoofr_kpt_dev(self, c0, f, wk, hgkp, hgkm, rhoe, nstate=None, ngroups=1, my_group=0, accumulate=False,
                       stream=None):
        """``cpb_rhoofr_kpt_dev``: one k-point of rhoofr_c (rhoofr_c_utils.mod.F90:117-178).
        Returns (ekin, rsum_g, rsum_r) contributions."""
        nstate, ld = self._kpt_args(c0, nstate)
        f = self._f_arg(f, nstate)
        self._g_arg(hgkp, "hgkp"), self._g_arg(hgkm, "hgkm")
        if _numel(rhoe) < self.nnr1:
            raise ValueError("rhoe too small")
        out = [C.c_double() for _ in range(3)]
        flags = _lib.CPB_RHO_ACCUMULATE if accumulate else 0
        self._check(self._L.cpb_rhoofr_kpt_dev(self._h, _ptr(c0), ld, nstate, f.ctypes.data, float(wk), _ptr(hgkp),
                                               _ptr(hgkm), ngroups, my_group, _ptr(rhoe),
                                               *[C.byref(o) for o in out], flags, _stream_ptr(stream)))
        return tuple(o.value for o in out)

    def vpsi_kpt_dev(self, c0, c2, f, hgkp, hgkm, vpot, nstate=None, ngroups=1, my_group=0, flags=0, stream=None):
        """``cpb_vpsi_kpt_dev``: vpsi's k-point branch for one k-point (vpsi_utils.mod.F90:562-625)."""
        nstate, ld = self._kpt_args(c0, nstate)
        f = self._f_arg(f, nstate)
        self._g_arg(hgkp, "hgkp"), self._g_arg(hgkm, "hgkm")
        if tuple(c2.shape) != tuple(c0.shape) or _numel(vpot) < self.nnr1:
            raise ValueError("c2 must have the shape of c0 and vpot nnr1 entries")
        self._check(self._L.cpb_vpsi_kpt_dev(self._h, _ptr(c0), _ptr(c2), ld, nstate, f.ctypes.data, _ptr(hgkp),
                                             _ptr(hgkm), _ptr(vpot), ngroups, my_group, flags, _stream_ptr(stream)))
        return c2

    # -- meta-GGA (cntl%ttau): gk is the Fortran gk(3, ngw) = a C-order (ngw, 3) array ---------------
    def tauofr_dev(self, c0, f, gk, tau, nsup=-1, nstate=None, ngroups=1, my_group=0, stream=None):
        """``cpb_tauofr_dev``: tau (nnr1,) or (2, nnr1) with LSD, zeroed and written
        (tauofr_utils.mod.F90:42-111)."""
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        need = (2 if nsup >= 0 else 1) * self.nnr1
        if (tau.numel() if _is_torch(tau) else tau.size) < need:
            raise ValueError("tau too small")
        self._g_arg(gk, "gk", 3)
        self._check(self._L.cpb_tauofr_dev(self._h, _ptr(c0), ld, nstate, f.ctypes.data, int(nsup), _ptr(gk), ngroups,
                                           my_group, _ptr(tau), 0, _stream_ptr(stream)))
        return tau

    def vtaupsi_dev(self, c0, c2, f, gk, vtau, nsup=-1, nstate=None, ngroups=1, my_group=0, stream=None):
        """``cpb_vtaupsi_dev``: c2 -= ... (vtaupsi_utils.mod.F90:38-165); vtau (nnr1,) or (2, nnr1)."""
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        self._g_arg(gk, "gk", 3)
        if tuple(c2.shape) != tuple(c0.shape) or _numel(vtau) < (2 if nsup >= 0 else 1) * self.nnr1:
            raise ValueError("c2 must have the shape of c0 and vtau nnr1 (2*nnr1 with LSD) entries")
        self._check(self._L.cpb_vtaupsi_dev(self._h, _ptr(c0), _ptr(c2), ld, nstate, f.ctypes.data, int(nsup), _ptr(gk),
                                            _ptr(vtau), ngroups, my_group, 0, _stream_ptr(stream)))
        return c2

    # -- Hartree-Fock exchange (hfx_old, Gamma point, no LSD, no screening) -------------------------
    def hfx_dev(self, dens_plan, c0, c2, f, scgx, pfl=0.25, nstate=None, stream=None):
        """``cpb_hfx_dev``: self = the wavefunction plan, ``dens_plan`` = the plan of the pair-density FFT set
        (same mesh), ``scgx`` its Coulomb kernel (device, dens_plan.ngw doubles).  c2 += C2_hfx; returns
        (ehfx, vhfx) (hfx_utils.mod.F90:80-965)."""
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        if tuple(c2.shape) != tuple(c0.shape):
            raise ValueError("c2 must have the shape of c0")
        if _numel(scgx) < dens_plan.ngw:
            raise ValueError("scgx needs one entry per vector of the pair-density set")
        e, v = C.c_double(), C.c_double()
        self._check(self._L.cpb_hfx_dev(self._h, dens_plan._h, _ptr(c0), _ptr(c2), ld, nstate, f.ctypes.data, _ptr(scgx),
                                        float(pfl), C.byref(e), C.byref(v), 0, _stream_ptr(stream)))
        return e.value, v.value

    def hfx(self, dens_plan, c0, c2, f, scgx, pfl=0.25, nstate=None):
        """``cpb_hfx`` (host arrays): c2 (in/out) += C2_hfx; returns (ehfx, vhfx)."""
        c0 = _as_host(c0, np.complex128)
        c2h = _as_host(c2, np.complex128)
        if c2h.shape != c0.shape:
            raise ValueError("c2 must have the shape of c0")
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        scgx = np.ascontiguousarray(scgx, dtype=np.float64)
        if scgx.size < dens_plan.ngw:
            raise ValueError("scgx needs one entry per vector of the pair-density set")
        e, v = C.c_double(), C.c_double()
        self._check(self._L.cpb_hfx(self._h, dens_plan._h, c0.ctypes.data, c2h.ctypes.data, ld, nstate, f.ctypes.data,
                                    scgx.ctypes.data, float(pfl), C.byref(e), C.byref(v), 0))
        return e.value, v.value

    # -- host-array forms of the k-point / meta-GGA entry points (the Fortran drop-in path) ----------
    def rhoofr_kpt(self, c0, f, wk, hgkp, hgkm, rhoe=None, nstate=None, ngroups=1, my_group=0, accumulate=False):
        c0 = _as_host(c0, np.complex128)
        nstate, ld = self._kpt_args(c0, nstate)
        f = self._f_arg(f, nstate)
        hgkp = self._g_arg(np.ascontiguousarray(hgkp, dtype=np.float64), 'hgkp')
        hgkm = self._g_arg(np.ascontiguousarray(hgkm, dtype=np.float64), 'hgkm')
        if rhoe is None:
            rhoe = np.zeros(self.nnr1, dtype=np.float64)
        rh = _as_host(rhoe, np.float64)
        out = [C.c_double() for _ in range(3)]
        flags = _lib.CPB_RHO_ACCUMULATE if accumulate else 0
        self._check(self._L.cpb_rhoofr_kpt(self._h, c0.ctypes.data, ld, nstate, f.ctypes.data, float(wk),
                                           hgkp.ctypes.data, hgkm.ctypes.data, ngroups, my_group, rh.ctypes.data,
                                           *[C.byref(o) for o in out], flags))
        return (rhoe, *[o.value for o in out])

    def vpsi_kpt(self, c0, c2, f, hgkp, hgkm, vpot, nstate=None, ngroups=1, my_group=0, flags=0):
        c0 = _as_host(c0, np.complex128)
        c2h = _as_host(c2, np.complex128)
        nstate, ld = self._kpt_args(c0, nstate)
        f = self._f_arg(f, nstate)
        hgkp = self._g_arg(np.ascontiguousarray(hgkp, dtype=np.float64), 'hgkp')
        hgkm = self._g_arg(np.ascontiguousarray(hgkm, dtype=np.float64), 'hgkm')
        v = _as_host(vpot, np.float64)
        self._check(self._L.cpb_vpsi_kpt(self._h, c0.ctypes.data, c2h.ctypes.data, ld, nstate, f.ctypes.data,
                                         hgkp.ctypes.data, hgkm.ctypes.data, v.ctypes.data, ngroups, my_group, flags))
        return c2

    def tauofr(self, c0, f, gk, tau=None, nsup=-1, nstate=None, ngroups=1, my_group=0):
        c0 = _as_host(c0, np.complex128)
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        gk = self._g_arg(np.ascontiguousarray(gk, dtype=np.float64), 'gk', 3)
        if tau is None:
            tau = np.empty((2 if nsup >= 0 else 1, self.nnr1), dtype=np.float64)
        th = _as_host(tau, np.float64)
        self._check(self._L.cpb_tauofr(self._h, c0.ctypes.data, ld, nstate, f.ctypes.data, int(nsup), gk.ctypes.data,
                                       ngroups, my_group, th.ctypes.data, 0))
        return tau

    def vtaupsi(self, c0, c2, f, gk, vtau, nsup=-1, nstate=None, ngroups=1, my_group=0):
        c0 = _as_host(c0, np.complex128)
        c2h = _as_host(c2, np.complex128)
        nstate, ld = self._c0_args(c0, nstate)
        f = self._f_arg(f, nstate)
        gk = self._g_arg(np.ascontiguousarray(gk, dtype=np.float64), 'gk', 3)
        vt = _as_host(vtau, np.float64)
        self._check(self._L.cpb_vtaupsi(self._h, c0.ctypes.data, c2h.ctypes.data, ld, nstate, f.ctypes.data, int(nsup),
                                        gk.ctypes.data, vt.ctypes.data, ngroups, my_group, 0))
        return c2

    # -- dense transforms on the density cutoff + local part of vofrho (plan built from nhg) --------
    # Arrays follow the package convention: Fortran (ld, nfields) = C-order (nfields, ld).
    def _dense_shapes(self, f, g):
        nf = 1 if g.ndim == 1 else int(g.shape[0])
        ld = int(g.shape[-1])
        n_f = f.numel() if _is_torch(f) else f.size
        if ld < self.ngw or n_f < nf * self.nnr1:
            raise ValueError("array shapes inconsistent with the plan (ngw, nnr1)")
        return nf, ld

    def dense_fwfft_dev(self, f, g, stream=None):
        """``cpb_dense_fwfft_dev``: f float64 CUDA (nfields, nnr1) -> g complex128 CUDA (nfields, ld):
        fwfftn(v,.FALSE.) + gather through nzh (vofrhoa_utils.mod.F90:88-95)."""
        nf, ld = self._dense_shapes(f, g)
        self._check(self._L.cpb_dense_fwfft_dev(self._h, _ptr(f), nf, _ptr(g), ld, _stream_ptr(stream)))
        return g

    def dense_invfft_dev(self, g, f, accumulate=False, stream=None):
        """``cpb_dense_invfft_dev``: g (nfields, ld) -> f (nfields, nnr1) = REAL/AIMAG of
        invfftn(v,.FALSE.) of the scattered coefficients (vofrhob_utils.mod.F90:155-173)."""
        nf, ld = self._dense_shapes(f, g)
        flags = _lib.CPB_DENSE_ACCUMULATE if accumulate else 0
        self._check(self._L.cpb_dense_invfft_dev(self._h, _ptr(g), ld, nf, _ptr(f), flags, _stream_ptr(stream)))
        return f

    @staticmethod
    def _ener_dict(e):
        return dict(eh=complex(e[0], e[1]), ei=complex(e[2], e[3]), ee=complex(e[4], e[5]),
                    eps=complex(e[6], e[7]), vploc=e[8])

    def vofrho_local_dev(self, rhoe, scg, eivps, eirop, v, rhog=None, vtemp=None, stream=None):
        """``cpb_vofrho_local_dev`` (CUDA tensors): rhoe -> rhog -> ppener -> v(r); v may be rhoe.
        Returns dict(eh, ei, ee, eps (complex), vploc) (ppener_utils.mod.F90:23-108)."""
        for a in (scg, eivps, eirop):
            if a.numel() < self.ngw:
                raise ValueError("scg / eivps / eirop need nhg entries")
        if rhoe.numel() < self.nnr1 or v.numel() < self.nnr1:
            raise ValueError("rhoe / v too small")
        e = (C.c_double * 9)()
        self._check(self._L.cpb_vofrho_local_dev(self._h, _ptr(rhoe), _ptr(scg), _ptr(eivps), _ptr(eirop),
                                                 _ptr(rhog), _ptr(vtemp), _ptr(v), e, _stream_ptr(stream)))
        return self._ener_dict(e)

    def vofrho_local(self, rhoe, scg, eivps, eirop, v=None, rhog=None, vtemp=None):
        """``cpb_vofrho_local`` (host arrays).  Returns (v, energies dict)."""
        rh = _as_host(rhoe, np.float64)
        scg = np.ascontiguousarray(scg, dtype=np.float64)
        eivps = np.ascontiguousarray(eivps, dtype=np.complex128)
        eirop = np.ascontiguousarray(eirop, dtype=np.complex128)
        if min(scg.size, eivps.size, eirop.size) < self.ngw or rh.size < self.nnr1:
            raise ValueError("array shapes inconsistent with the plan (ngw, nnr1)")
        if v is None:
            v = np.empty(self.nnr1, dtype=np.float64)
        vh = _as_host(v, np.float64)
        e = (C.c_double * 9)()
        self._check(self._L.cpb_vofrho_local(self._h, rh.ctypes.data, scg.ctypes.data, eivps.ctypes.data,
                                             eirop.ctypes.data, _ptr(rhog), _ptr(vtemp), vh.ctypes.data, e))
        return v, self._ener_dict(e)


def _as_host(a, dtype):
    """numpy view of a host array (numpy, or a CPU/pinned torch tensor) without copying."""
    if isinstance(a, np.ndarray):
        if a.dtype != dtype or not a.flags.c_contiguous:
            raise ValueError(f"expected a C-contiguous {np.dtype(dtype).name} array")
        return a
    if _is_torch(a):
        if a.is_cuda:
            raise ValueError("host entry point called with a CUDA tensor; use the *_dev variant")
        return a.numpy()
    raise TypeError("expected numpy array or torch tensor")


def _stream_ptr(stream):
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except Exception:
            pass
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)


# ---------------------------------------------------------------------------------------------
# mirror of the reference's module-global state + subroutine signatures
# ---------------------------------------------------------------------------------------------

@dataclass
class CpmdContext:
    """The module globals the two subroutines read, gathered in one object.

    spar%nr1s.. -> ``nr``; fpar%kr1.. -> ``kr``; ncpw%ngw -> ``ngw``; cppt inyh/hg; parm%tpiba2,
    parm%omega; crge%f(:,1) -> ``f`` (rhoofr's occupations); parai%cp_nogrp / cp_inter_me ->
    ``cp_nogrp`` / ``cp_inter_me``; the variant flags that the GPU path does not implement.
    """
    nr: tuple
    inyh: np.ndarray
    hg: np.ndarray
    tpiba2: float = 1.0
    omega: float = 1.0
    f: np.ndarray = None              # crge%f(:,1)
    cp_nogrp: int = 1                 # parai%cp_nogrp
    cp_inter_me: int = 0              # parai%cp_inter_me
    device: int = 0
    max_batch: int = 32
    # variant switches (must all be off; otherwise the shim falls back to the original routine)
    tkpnt: bool = False               # tkpts%tkpnt
    tlsd: bool = False                # cntl%tlsd
    tlse: bool = False                # lspin2%tlse
    ttau: bool = False                # cntl%ttau
    tdg: bool = False                 # tdgcomm%tdg
    rsactive: bool = False
    nsup: int = 0                     # spin_mod%nsup (number of alpha states, used with tlsd)
    tksham: bool = False              # cntl%tksham
    akin: float = 0.0                 # prcp_com%akin
    nogrp: int = 1                    # group%nogrp (old task groups)
    delta: float = 1.0e-6             # rhoofr charge tolerance (rhoofr_utils.mod.F90:141)
    plan: Plan = field(default=None, repr=False)
    _cdll: object = field(default=None, repr=False)
    # outputs the reference stores in globals
    ekin: float = 0.0                 # ener_com%ekin
    csumg: float = 0.0                # chrg%csumg
    csumr: float = 0.0                # chrg%csumr
    csums: float = 0.0                # chrg%csums    (LSD)
    csumsabs: float = 0.0             # chrg%csumsabs (LSD)

    def __post_init__(self):
        if self.plan is None:
            self.plan = Plan(self.nr, self.inyh, self.hg, self.tpiba2, self.omega, device=self.device,
                             max_batch=self.max_batch, _cdll=self._cdll)
        self.kr = self.plan.kr
        self.ngw = self.plan.ngw
        self.nnr1 = self.plan.nnr1

    def _check_variant(self, proc):
        if self.nogrp > 1:
            raise StopGM(proc, "OLD TASK GROUPS NOT SUPPORTED ANYMORE ")  # vpsi_utils.mod.F90:173-175
        for flag, name in ((self.tkpnt, "k-points"), (self.tlse, "LSE"),
                           (self.ttau, "meta-GGA tau"), (self.tdg, "double grid"),
                           (self.rsactive, "REAL SPACE WFN KEEP"), (self.akin > 1.0e-10, "AKIN")):
            if flag:
                raise StopGM(proc, f"{name} variant is not implemented on the GPU path")

    def rhoofr(self, c0, rhoe, psi, nstate):
        """``SUBROUTINE rhoofr(c0,rhoe,psi,nstate)`` (rhoofr_utils.mod.F90:122-137).
        ``psi`` is the caller's scratch array; unused (the library owns its work space).
        Sets ``ekin``, ``csumg``, ``csumr`` like the reference sets ener_com%ekin, chrg%csum*.
        With cp_nogrp > 1 ``rhoe``/sums are the group's partial results: the caller performs
        cp_grp_redist (see cpmd_b200.dist)."""
        proc = "rhoofr"
        self._check_variant(proc)
        if self.f is None:
            raise StopGM(proc, "occupation numbers crge%f not set")
        dev = _is_torch(c0) and c0.is_cuda
        rh = rhoe if rhoe.ndim == 1 else rhoe.reshape(-1)
        if self.tlsd:
            # rhoe(nnr1, nlsd=2): a (2, nnr1) array here
            if dev:
                ekin, rg, rr, cs, ca = self.plan.rhoofr_lsd_dev(c0, self.f, self.nsup, rh, nstate, self.cp_nogrp,
                                                                self.cp_inter_me)
            else:
                _, ekin, rg, rr, cs, ca = self.plan.rhoofr_lsd(c0, self.f, self.nsup, rh, nstate, self.cp_nogrp,
                                                               self.cp_inter_me)
            self.csums, self.csumsabs = cs, ca
        elif dev:
            ekin, rg, rr = self.plan.rhoofr_dev(c0, self.f, rh, nstate, self.cp_nogrp, self.cp_inter_me)
        else:
            _, ekin, rg, rr = self.plan.rhoofr(c0, self.f, rh, nstate, self.cp_nogrp, self.cp_inter_me)
        self.ekin, self.csumg, self.csumr = ekin, rg, rr
        if self.cp_nogrp == 1 and abs(rr - rg) > self.delta:
            raise StopGM(proc, "TOTAL DENSITY SUMS ARE NOT EQUAL")  # :625-635

    def vpsi(self, c0, c2, f, vpot, psi, nstate, ikind=1, ispin=1, redist_c2=False):
        """``SUBROUTINE vpsi(c0,c2,f,vpot,psi,nstate,ikind,ispin,redist_c2)``
        (vpsi_utils.mod.F90:120-135).  c2 += -(f/2) [tpiba2 hg c0 + 2 FFT(V psi)] for the
        group's states.  ``redist_c2`` is the caller's job here (cpmd_b200.dist.redist_c2)."""
        proc = "vpsi"
        self._check_variant(proc)
        if ikind != 1:
            raise StopGM(proc, "k-points (ikind>1) not implemented on the GPU path")
        flags = _lib.CPB_VPSI_TKSHAM if self.tksham else 0
        v = vpot if vpot.ndim == 1 else vpot.reshape(-1)
        if self.tlsd and ispin == 2:       # vpsi_utils.mod.F90:450: cntl%tlsd .AND. ispin == 2
            if _is_torch(c0) and c0.is_cuda:
                self.plan.vpsi_lsd_dev(c0, c2, f, self.nsup, v, nstate, self.cp_nogrp, self.cp_inter_me, flags)
            else:
                self.plan.vpsi_lsd(c0, c2, f, self.nsup, v, nstate, self.cp_nogrp, self.cp_inter_me, flags)
            return
        if ispin != 1:
            raise StopGM(proc, "ispin=2 without cntl%tlsd")
        if _is_torch(c0) and c0.is_cuda:
            self.plan.vpsi_dev(c0, c2, f, v, nstate, self.cp_nogrp, self.cp_inter_me, flags)
        else:
            self.plan.vpsi(c0, c2, f, v, nstate, self.cp_nogrp, self.cp_inter_me, flags)
