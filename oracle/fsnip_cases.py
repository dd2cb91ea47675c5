"""The reference's own Fortran statements (executed by oracle/fsnip.py) assembled into the non-transform parts of
``rhoofr`` and ``vpsi``: pairing loops, occupation rules, +-G unpack with the kinetic term, density accumulation,
``kin_energy`` / ``dotp``.  The transforms between them (``set_psi_*``, ``invfftn``, ``fwfftn``) come from the oracle:
those are pinned by the reference's helper kernels in oracle/_ref (tests/test_oracle_ref.py).  TEST INFRASTRUCTURE:
used by tests/test_fsnip_pin.py and tools/make_golden_fsnip.py only."""
from __future__ import annotations

import numpy as np

from oracle import cpmd_oracle as orc
from oracle import fsnip
from oracle.fsnip import FArr, ns


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
        if is2 > nstate:
            psi = orc.set_psi_1_state_g(geo, c0[is1 - 1])
        else:
            psi = orc.set_psi_2_states_g(geo, c0[is1 - 1], c0[is2 - 1])
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
        if is2 > nstate:
            psi = orc.set_psi_1_state_g(geo, c0[is1 - 1])
        else:
            psi = orc.set_psi_2_states_g(geo, c0[is1 - 1], c0[is2 - 1])
        psi = orc.fwfftn_sparse(geo, vpot * orc.invfftn_sparse(geo, psi))
        env = dict(f=FArr(np.asarray(f, float)), is1=is1, is2=is2, nostat=nstate, cntl=ns(tksham=bool(tksham)),
                   prcp_com=ns(akin=0.0, gskin=1.0, gckin=0.0, gakin=0.0), psi_p=FArr(psi), nzhs=FArr(geo.nzhs),
                   indzs=FArr(geo.indzs), jgw=geo.ngw, hg=FArr(geo.hg), parm=ns(tpiba2=tpiba2), c0=c0f,
                   C2_vpsi=FArr(c2v))
        fsnip.run("vpsi_utils.mod.F90", 627, 672, env)
    return c2 + c2v.T
