"""The reference's own Fortran statements (executed by oracle/fsnip.py) assembled into the non-transform parts of
``rhoofr`` and ``vpsi``: pairing loops, occupation rules, +-G unpack with the kinetic term, density accumulation,
``kin_energy`` / ``dotp``, ``set_psi_*``.  The transforms between them (``invfftn``, ``fwfftn``) come from the oracle:
those are pinned by the reference's helper kernels in oracle/_ref (tests/test_oracle_ref.py).  TEST INFRASTRUCTURE:
used by tests/test_fsnip_pin.py and tools/make_golden_fsnip.py only."""
from __future__ import annotations

import numpy as np

from oracle import cpmd_oracle as orc
from oracle import fsnip
from oracle.fsnip import FArr, ns


def set_psi(geo, c1, c2=None):
    """Ray storage psi of one state (c2 None: set_psi_1_state_g with alpha = 1, state_utils.mod.F90:142-166) or of a
    packed pair (set_psi_2_states_g, :183-187), from the reference's statements."""
    psi = np.zeros(geo.kr[0] * geo.nrays, complex)
    env = dict(jgw=geo.ngw, psi=FArr(psi), nzfs=FArr(geo.nzhs), inzs=FArr(geo.indzs), geq0=bool(geo.geq0), c1=FArr(c1),
               uimag=1j)
    if c2 is None:
        env.update(alpha=1.0 + 0.0j, zone=1.0 + 0.0j)
        fsnip.run("state_utils.mod.F90", 142, 166, env)
    else:
        env.update(c2=FArr(c2))
        fsnip.run("state_utils.mod.F90", 183, 187, env)
    return psi


def part_1d_functions():
    """part_1d_nbr_el_in_blk / part_1d_get_el_in_blk from the statements of part_1d.mod.F90:31-34 and :51-53."""
    def nbr_el(n_elem, proc, nproc):
        env = fsnip.run("part_1d.mod.F90", 31, 34, dict(n_elem=int(n_elem), proc=int(proc), nproc=int(nproc)))
        return int(env["part_1d_nbr_el_in_blk"])

    def get_el(i_elem, n_elem, proc, nproc):
        env = fsnip.run("part_1d.mod.F90", 51, 53,
                        dict(i_elem=int(i_elem), n_elem=int(n_elem), proc=int(proc), nproc=int(nproc)))
        return int(env["part_1d_get_el_in_blk"])
    return nbr_el, get_el


def pair_loop(which, nstate, group, ngroups):
    """(is1, is2) (1-based, is2 = nstate+1 for a trailing single state) from the loop headers
    rhoofr_utils.mod.F90:306-310 (which='rhoofr') or vpsi_utils.mod.F90:376-383 (which='vpsi', njump = 2)."""
    nbr_el, get_el = part_1d_functions()
    out = []
    env = dict(part_1d_nbr_el_in_blk=nbr_el, part_1d_get_el_in_blk=get_el,
               parai=ns(cp_inter_me=group, cp_nogrp=ngroups), nstate=nstate, nostat=nstate, njump=2,
               record=lambda a, b: out.append((int(a), int(b))))
    if which == "rhoofr":
        fsnip.run("rhoofr_utils.mod.F90", 306, 310, env, tail=["CALL record(is1,is2)", "ENDDO"])
    else:
        fsnip.run("vpsi_utils.mod.F90", 376, 383, env, tail=["CALL record(is1,is2)", "ENDDO"])
    return out


def dotp_fn(geq0):
    """dotp_c (dotp_utils.mod.F90:45-52) as a Python callable dotp(n, a, b)."""
    def dotp(n, a, b):
        env = fsnip.run("dotp_utils.mod.F90", 45, 52, dict(n=int(n), a=a, b=b, geq0=bool(geq0)))
        return env["dotp"]
    return dotp


def kin_energy(geo, c0, f, tpiba2):
    """kin_energy_utils.mod.F90:62-110 (akin = 0): returns (ekin, rsum)."""
    nstate = c0.shape[0]
    env = dict(nstate=nstate, crge=ns(f=FArr(np.asarray(f, float).reshape(nstate, 1))), ncpw=ns(ngw=geo.ngw),
               c0=FArr(np.ascontiguousarray(c0.T)), hg=FArr(geo.hg), dotp=dotp_fn(True), geq0=True,
               prcp_com=ns(akin=0.0, gskin=1.0, gckin=0.0, gakin=0.0), deltakin=1.0e-10, parm=ns(tpiba2=tpiba2),
               ener_com=ns(ekin=0.0), ngw=geo.ngw)
    fsnip.run("kin_energy_utils.mod.F90", 62, 110, env)
    return float(env["ener_com"].ekin), float(env["rsum"])


def rhoofr(geo, c0, f, omega, tpiba2, group=0, ngroups=1):
    """rhoofr with the reference's statements for the pair loop (:306-310), the coefficients (:369-374) and
    build_density_sum (density_utils.mod.F90:77-80); transforms from the oracle."""
    nstate = c0.shape[0]
    rho = np.zeros(geo.nnr1)
    fa = FArr(np.asarray(f, float).reshape(nstate, 1))
    for is1, is2 in pair_loop("rhoofr", nstate, group, ngroups):
        if f[is1 - 1] == 0.0 and (is2 > nstate or f[is2 - 1] == 0.0):
            continue                                                       # tfcal, rhoofr_utils.mod.F90:312-316
        psi = set_psi(geo, c0[is1 - 1], None if is2 > nstate else c0[is2 - 1])
        psi = orc.invfftn_sparse(geo, psi)
        env = dict(crge=ns(f=fa), parm=ns(omega=omega), is1=is1, is2=is2, nstate=nstate)
        fsnip.run("rhoofr_utils.mod.F90", 369, 374, env)
        env2 = dict(alpha_real=env["coef3"], alpha_imag=env["coef4"], psi=FArr(psi), rho=FArr(rho), n=geo.nnr1)
        fsnip.run("density_utils.mod.F90", 77, 80, env2)
    return rho


def vpsi(geo, c0, c2, f, vpot, tpiba2, group=0, ngroups=1, tksham=False):
    """vpsi with the reference's statements for the pair loop (:376-383), the occupation rules and the +-G
    unpack with the kinetic term (:627-672, akin = 0); transforms from the oracle; c2 += C2_vpsi (:717)."""
    nstate = c0.shape[0]
    c2v = np.zeros((geo.ngw, nstate), complex)                             # column-major C2_vpsi(ngw, nstate)
    c0f = FArr(np.ascontiguousarray(c0.T))
    for is1, is2 in pair_loop("vpsi", nstate, group, ngroups):
        psi = set_psi(geo, c0[is1 - 1], None if is2 > nstate else c0[is2 - 1])
        psi = orc.fwfftn_sparse(geo, vpot * orc.invfftn_sparse(geo, psi))
        env = dict(f=FArr(np.asarray(f, float)), is1=is1, is2=is2, nostat=nstate, cntl=ns(tksham=bool(tksham)),
                   prcp_com=ns(akin=0.0, gskin=1.0, gckin=0.0, gakin=0.0), psi_p=FArr(psi), nzhs=FArr(geo.nzhs),
                   indzs=FArr(geo.indzs), jgw=geo.ngw, hg=FArr(geo.hg), parm=ns(tpiba2=tpiba2), c0=c0f,
                   C2_vpsi=FArr(c2v))
        fsnip.run("vpsi_utils.mod.F90", 627, 672, env)
    return c2 + c2v.T


# ---------------------------------------------------------------------------------------------------------------
# k-points: one k-point of rhoofr_c (rhoofr_c_utils.mod.F90:117-178) and the k-point branch of vpsi's unpack
# (vpsi_utils.mod.F90:562-625); set_psi_1_state_g_kpts from state_utils.mod.F90:202-222
# ---------------------------------------------------------------------------------------------------------------
def _set_psi_kpts_fn(geo):
    def set_psi_1_state_g_kpts(alpha, c1, psi):
        fsnip.run("state_utils.mod.F90", 202, 222,
                  dict(alpha=alpha, zone=1.0 + 0.0j, c1=c1, psi=psi, ncpw=ns(ngw=geo.ngw), nzhs=FArr(geo.nzhs),
                       indzs=FArr(geo.indzs), geq0=bool(geo.geq0)))
    return set_psi_1_state_g_kpts


def _build_density_sum(alpha_real, alpha_imag, psi, rho, n):
    fsnip.run("density_utils.mod.F90", 77, 80, dict(alpha_real=alpha_real, alpha_imag=alpha_imag, psi=psi, rho=rho, n=int(n)))


def rhoofr_kpt(geo, c0, f, wk, hgkp, hgkm, omega, tpiba2, group=0, ngroups=1):
    """One k-point of rhoofr_c: the reference's statements :117-178 and :182, transforms (invfftn) from the oracle."""
    nstate = c0.shape[0]
    nbr_el, get_el = part_1d_functions()
    npsi = max(geo.nnr1, geo.kr[0] * geo.nrays)
    psi = FArr(np.zeros(npsi, complex))
    rhoe = np.zeros((geo.nnr1, 1), order="F")

    def invfftn(p, sparse, comm):
        r = orc.invfftn_sparse(geo, p.a[:geo.kr[0] * geo.nrays].copy())
        p.a[:] = 0.0
        p.a[:geo.nnr1] = r

    def zeroing(p):
        p.a[:] = 0.0

    c0f = np.asfortranarray(c0.T.reshape(2 * geo.ngw, nstate, 1))
    env = dict(nstate=nstate, ikind=1, ikk=1, crge=ns(f=FArr(np.asarray(f, float).reshape(nstate, 1))), wk=FArr(np.array([wk])),
               nkpt=ns(ngwk=2 * geo.ngw), ncpw=ns(ngw=geo.ngw), c0=FArr(c0f), hgkp=FArr(hgkp.reshape(-1, 1)),
               hgkm=FArr(hgkm.reshape(-1, 1)), prcp_com=ns(akin=0.0, gskin=1.0, gckin=0.0, gakin=0.0), deltakin=1.0e-10,
               rsum=0.0, xkin=0.0, part_1d_nbr_el_in_blk=nbr_el, part_1d_get_el_in_blk=get_el,
               parai=ns(cp_inter_me=group, cp_nogrp=ngroups, allgrp=0), rsactive=False, psi=psi, zone=1.0 + 0.0j,
               zeroing=zeroing, set_psi_1_state_g_kpts=_set_psi_kpts_fn(geo), invfftn=invfftn, cntl=ns(tlsd=False),
               parm=ns(omega=omega, tpiba2=tpiba2), build_density_sum=_build_density_sum, rhoe=FArr(rhoe),
               fpar=ns(nnr1=geo.nnr1), ener_com=ns(ekin=0.0), spin_mod=ns(nsup=nstate), maxstates=0)
    fsnip.run("rhoofr_c_utils.mod.F90", 117, 178, env)
    fsnip.run("rhoofr_c_utils.mod.F90", 182, 182, env)
    return dict(rhoe=rhoe[:, 0].copy(), ekin=float(env["ener_com"].ekin), rsum_g=float(env["rsum"]))


def vpsi_kpt(geo, c0, c2, f, hgkp, hgkm, vpot, tpiba2, group=0, ngroups=1):
    """vpsi with tkpts%tkpnt for one k-point: the reference's unpack statements :562-625, transforms from the oracle."""
    nstate = c0.shape[0]
    ngw = geo.ngw
    nbr_el, get_el = part_1d_functions()
    c2v = np.zeros((2 * ngw, nstate), complex, order="F")
    c0f = FArr(np.asfortranarray(c0.T))
    setpsi = _set_psi_kpts_fn(geo)
    for i in range(1, nbr_el(nstate, group, ngroups) + 1):
        is1 = get_el(i, nstate, group, ngroups)
        psi = FArr(np.zeros(geo.kr[0] * geo.nrays, complex))
        setpsi(1.0 + 0.0j, c0f(slice(None), is1), psi)
        out = orc.fwfftn_sparse(geo, vpot * orc.invfftn_sparse(geo, psi.a))
        env = dict(tkpts=ns(tkpnt=True), f=FArr(np.asarray(f, float)), is1=is1, cntl=ns(tgaugep=False, tgaugef=False),
                   ncpw=ns(ngw=ngw), psi_p=FArr(out), nzhs=FArr(geo.nzhs), indzs=FArr(geo.indzs), C2_vpsi=FArr(c2v),
                   parm=ns(tpiba2=tpiba2), hgkp=FArr(hgkp.reshape(-1, 1)), hgkm=FArr(hgkm.reshape(-1, 1)), ikind=1, c0=c0f,
                   geq0=bool(geo.geq0))
        fsnip.run("vpsi_utils.mod.F90", 562, 625, env, tail=["ENDIF"])
    return c2 + c2v.T


def ppener(geo_d, rhog, scg, eivps, eirop, geq0=True):
    """ppener (ppener_utils.mod.F90:58-104): (eh, ei, ee, eps, vploc, vtemp) from the reference's statements; v(nzh(ig))
    is handed over as the coefficient list itself (nzh = identity)."""
    nhg = len(rhog)
    vtemp = np.zeros(nhg, complex)
    env = dict(geq0=bool(geq0), eivps=FArr(eivps), eirop=FArr(eirop), v=FArr(rhog), nzh=FArr(np.arange(1, nhg + 1)),
               scg=FArr(scg), ncpw=ns(nhg=nhg), vtemp=FArr(vtemp))
    fsnip.run("ppener_utils.mod.F90", 58, 104, env)
    return env["eh"], env["ei"], env["ee"], env["eps"], env["vploc"], vtemp


# ---------------------------------------------------------------------------------------------------------------
# meta-GGA: tauofr (tauofr_utils.mod.F90:82-102 with dpsisc :121-135 and tauadd :148-173) and vtaupsi
# (vtaupsi_utils.mod.F90:63-89 with taupot :103-127 and ftauadd :142-163)
# ---------------------------------------------------------------------------------------------------------------
def _tau_common(geo, c0, gk, nstate, group, ngroups, nsup):
    nbr_el, get_el = part_1d_functions()
    nray = geo.kr[0] * geo.nrays
    npsi = max(geo.nnr1, nray)
    psi = FArr(np.zeros(npsi, complex))
    base = dict(ncpw=ns(ngw=geo.ngw), nzhs=FArr(geo.nzhs), indzs=FArr(geo.indzs), gk=FArr(np.asfortranarray(gk.T)),
                uimag=1j, geq0=bool(geo.geq0), fpar=ns(nnr1=geo.nnr1), spin_mod=ns(nsup=nstate if nsup is None else nsup),
                cntl=ns(tlsd=nsup is not None))

    def zeroing(p):
        p.a[:] = 0.0

    def invfftn(p, sparse, comm):
        r = orc.invfftn_sparse(geo, p.a[:nray].copy())
        p.a[:] = 0.0
        p.a[:geo.nnr1] = r

    def fwfftn(p, sparse, comm):
        r = orc.fwfftn_sparse(geo, p.a[:geo.nnr1].copy())
        p.a[:] = 0.0
        p.a[:nray] = r

    def dpsisc(c0_, psi_, k, is1, is2, nstate_):
        fsnip.run("tauofr_utils.mod.F90", 121, 135, dict(base, c0=c0_, psi=psi_, k=int(k), is1=int(is1), is2=int(is2), nstate=nstate_))

    env = dict(base, nstate=nstate, part_1d_nbr_el_in_blk=nbr_el, part_1d_get_el_in_blk=get_el,
               parai=ns(cp_inter_me=group, cp_nogrp=ngroups, allgrp=0), psi=psi, c0=FArr(np.asfortranarray(c0.T)),
               zeroing=zeroing, invfftn=invfftn, fwfftn=fwfftn, dpsisc=dpsisc)
    return env, base


def tauofr(geo, c0, f, gk, omega, tpiba2, nsup=None, group=0, ngroups=1):
    """tau (nlsd, nnr1) of the group's block from the reference's statements; gk: (ngw, 3) like the oracle's."""
    nstate = c0.shape[0]
    nlsd = 1 if nsup is None else 2
    tau = np.zeros((geo.nnr1, nlsd), order="F")
    env, base = _tau_common(geo, c0, gk, nstate, group, ngroups, nsup)

    def tauadd(psi_, tau_, is1, is2, nstate_):
        fsnip.run("tauofr_utils.mod.F90", 148, 173,
                  dict(base, psi=psi_, tau=tau_, is1=int(is1), is2=int(is2), nstate=nstate_, parm=ns(tpiba2=tpiba2, omega=omega),
                       crge=ns(f=FArr(np.asarray(f, float).reshape(nstate, 1)))))

    env.update(tauadd=tauadd, tau=FArr(tau))
    fsnip.run("tauofr_utils.mod.F90", 82, 102, env)
    return np.ascontiguousarray(tau.T)


def vtaupsi(geo, c0, c2, f, gk, vtau, tpiba2, nsup=None, group=0, ngroups=1):
    """c2 updated by vtaupsi from the reference's statements; vtau: (ispin, nnr1)."""
    nstate = c0.shape[0]
    vt = np.atleast_2d(vtau)
    ispin = vt.shape[0]
    out = np.asfortranarray(c2.T.copy())
    env, base = _tau_common(geo, c0, gk, nstate, group, ngroups, nsup)

    def taupot(vpot_, psi_, is1, is2, ispin_):
        fsnip.run("vtaupsi_utils.mod.F90", 103, 127, dict(base, vpot=vpot_, psi=psi_, is1=int(is1), is2=int(is2), ispin=int(ispin_)))

    def ftauadd(c2_, psi_, f_, k, is1, is2, nstate_):
        fsnip.run("vtaupsi_utils.mod.F90", 142, 163,
                  dict(base, c2=c2_, psi=psi_, f=f_, k=int(k), is1=int(is1), is2=int(is2), nstate=nstate_, parm=ns(tpiba2=tpiba2)))

    env.update(taupot=taupot, ftauadd=ftauadd, vpot=FArr(np.asfortranarray(vt.T)), c2=FArr(out),
               f=FArr(np.asarray(f, float)), ispin=ispin)
    fsnip.run("vtaupsi_utils.mod.F90", 63, 89, env)
    return np.ascontiguousarray(out.T)


# ---------------------------------------------------------------------------------------------------------------
# exact exchange: hfxab (hfx_utils.mod.F90:1050-1106) and hfxaa (:1216-1256)
# ---------------------------------------------------------------------------------------------------------------
def _hfx_env(geo_w, geo_d, scgx, omega):
    nray_d, nray_w = geo_d.kr[0] * geo_d.nrays, geo_w.kr[0] * geo_w.nrays
    npsi = max(geo_w.nnr1, nray_d, nray_w)

    def zeroing(p):
        p.a[:] = 0.0

    def dscal(n, alpha, p, inc):
        v = p.a.view(np.float64)
        v[:int(n)] *= alpha

    def fwfftn(p, sparse, comm):
        r = orc.fwfftn_sparse(geo_w, p.a[:geo_w.nnr1].copy()) if sparse else orc.fwfftn_dense(geo_d, p.a[:geo_d.nnr1].copy())
        p.a[:] = 0.0
        p.a[:len(r)] = r

    def invfftn(p, sparse, comm):
        assert not sparse
        r = orc.invfftn_dense(geo_d, p.a[:nray_d].copy())
        p.a[:] = 0.0
        p.a[:len(r)] = r

    return dict(llr1=geo_w.nnr1, jhg=geo_d.ngw, jgw=geo_w.ngw, parm=ns(omega=omega), parai=ns(allgrp=0), scgx=FArr(scgx),
                nzff=FArr(geo_d.nzhs), inzf=FArr(geo_d.indzs), nzfs=FArr(geo_w.nzhs), inzs=FArr(geo_w.indzs),
                geq0=bool(geo_w.geq0), uimag=1j, zeroing=zeroing, dscal=dscal, fwfftn=fwfftn, invfftn=invfftn,
                psic=FArr(np.zeros(npsi, complex)), vpotg=FArr(np.zeros(geo_d.ngw, complex)),
                vpotr=FArr(np.zeros(geo_w.nnr1)))


def hfxab(geo_w, geo_d, psia, psib, iran, pf, scgx, omega):
    """(ehfx, dc2a, dc2b) of one pair from the reference's statements (transforms from the oracle)."""
    env = _hfx_env(geo_w, geo_d, scgx, omega)
    c2a, c2b = np.zeros(geo_w.ngw, complex), np.zeros(geo_w.ngw, complex)
    env.update(psia=FArr(psia), psib=FArr(psib), iran=int(iran), pf=pf, c2a=FArr(c2a), c2b=FArr(c2b))
    fsnip.run("hfx_utils.mod.F90", 1050, 1106, env)
    return float(env["ehfx"]), c2a, c2b


def hfxaa(geo_w, geo_d, psia, pf, scgx, omega):
    """(ehfx, dc2a) of the diagonal term from the reference's statements."""
    env = _hfx_env(geo_w, geo_d, scgx, omega)
    c2a = np.zeros(geo_w.ngw, complex)
    env.update(psia=FArr(psia), pf=pf, c2a=FArr(c2a))
    fsnip.run("hfx_utils.mod.F90", 1216, 1256, env)
    return float(env["ehfx"]), c2a


# ---------------------------------------------------------------------------------------------------------------
# LSD: rhoofr with the spin-resolved accumulation (rhoofr_utils.mod.F90:369-385, build_density_real / _imag
# density_utils.mod.F90:31-33 / 54-56) and vpsi with the spin-resolved potential (vpsi_utils.mod.F90:450-482)
# ---------------------------------------------------------------------------------------------------------------
def rhoofr_lsd(geo, c0, f, omega, tpiba2, nsup, group=0, ngroups=1):
    """(2, nnr1) channel densities [alpha, beta] of the group's block before the alpha+beta step (:543-559)."""
    nstate = c0.shape[0]
    rho = np.zeros((geo.nnr1, 2), order="F")
    fa = FArr(np.asarray(f, float).reshape(nstate, 1))

    def build_density_real(alpha, psi, rho_, n):
        fsnip.run("density_utils.mod.F90", 31, 33, dict(alpha=alpha, psi=psi, rho=rho_, n=int(n)))

    def build_density_imag(alpha, psi, rho_, n):
        fsnip.run("density_utils.mod.F90", 54, 56, dict(alpha=alpha, psi=psi, rho=rho_, n=int(n)))

    for is1, is2 in pair_loop("rhoofr", nstate, group, ngroups):
        if f[is1 - 1] == 0.0 and (is2 > nstate or f[is2 - 1] == 0.0):
            continue
        psi = set_psi(geo, c0[is1 - 1], None if is2 > nstate else c0[is2 - 1])
        psi = orc.invfftn_sparse(geo, psi)
        env = dict(crge=ns(f=fa), parm=ns(omega=omega), is1=is1, is2=is2, nstate=nstate, cntl=ns(tlsd=True),
                   spin_mod=ns(nsup=nsup), psi_p=FArr(psi), rhoe_p=FArr(rho), llr1=geo.nnr1,
                   build_density_sum=_build_density_sum, build_density_real=build_density_real,
                   build_density_imag=build_density_imag)
        fsnip.run("rhoofr_utils.mod.F90", 369, 385, env, tail=["ENDIF"])
    return np.ascontiguousarray(rho.T)


def vpsi_lsd(geo, c0, c2, f, vpot2, tpiba2, nsup, group=0, ngroups=1, tksham=False):
    """vpsi with cntl%tlsd, ispin = 2: the reference's potential application (:450-482) and unpack (:627-672)."""
    nstate = c0.shape[0]
    c2v = np.zeros((geo.ngw, nstate), complex)
    c0f = FArr(np.ascontiguousarray(c0.T))
    vdg = np.asfortranarray(np.asarray(vpot2).T)                             # vpotdg(nnr1, 2)
    vx = vdg.reshape(-1, order="F")                                          # vpotx: the same storage, 1-D
    for is1, is2 in pair_loop("vpsi", nstate, group, ngroups):
        psi = set_psi(geo, c0[is1 - 1], None if is2 > nstate else c0[is2 - 1])
        psi = orc.invfftn_sparse(geo, psi)
        envp = dict(cntl=ns(tlsd=True), ispin=2, is1=is1, spin_mod=ns(nsup=nsup), td_prop=ns(td_extpot=False),
                    nnrx=geo.nnr1, psi_p=FArr(psi), vpotx=FArr(vx), leadx=geo.nnr1, uimag=1j, vpotdg=FArr(vdg))
        fsnip.run("vpsi_utils.mod.F90", 450, 482, envp, tail=["ENDIF"])
        psi = orc.fwfftn_sparse(geo, psi)
        env = dict(f=FArr(np.asarray(f, float)), is1=is1, is2=is2, nostat=nstate, cntl=ns(tksham=bool(tksham)),
                   prcp_com=ns(akin=0.0, gskin=1.0, gckin=0.0, gakin=0.0), psi_p=FArr(psi), nzhs=FArr(geo.nzhs),
                   indzs=FArr(geo.indzs), jgw=geo.ngw, hg=FArr(geo.hg), parm=ns(tpiba2=tpiba2), c0=c0f,
                   C2_vpsi=FArr(c2v))
        fsnip.run("vpsi_utils.mod.F90", 627, 672, env)
    return c2 + c2v.T
