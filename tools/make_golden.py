#!/usr/bin/env python3
"""Generate tests/golden/*.npz: small input/output vectors of the hot path.

The reference (CPMD 4.3, Fortran + FFTW + MPI) cannot be built or imported in this container and
ships no fixtures for vpsi/rhoofr, so these vectors come from oracle/cpmd_oracle.py (the NumPy
restatement, itself pinned by tests/test_oracle.py's known-answer tests and cross-checked against
oracle/staged_oracle.c).  They freeze the oracle: any later change of its numerics shows up as a
diff against these files.  PARITY UNPINNED with respect to the reference's own tests.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpmd_oracle as orc  # noqa: E402

CASES = [
    # name, mesh, nstate, f_pattern, omega, tpiba2, (group, ngroups)
    ("n16_s5_mixed", (16, 16, 16), 5, "mixed", 1.7, 0.8, (0, 1)),
    ("n20_s4_all2", (20, 20, 20), 4, "all2", 1.0, 1.0, (0, 1)),
    ("n16x20x24_s3", (16, 20, 24), 3, "all2", 2.0, 1.1, (0, 1)),
    ("n16_s7_grp1of3", (16, 16, 16), 7, "mixed", 1.0, 1.0, (1, 3)),
]

VOFRHO_CASES = [
    # name, mesh, omega, tpiba2
    ("n16", (16, 16, 16), 1.3, 0.9),
    ("n16x20x24", (16, 20, 24), 2.0, 1.1),
]

KPT_CASES = [
    # name, mesh, nstate, kvec, wk, omega, tpiba2
    ("n16_s5", (16, 16, 16), 5, (0.25, 0.1, -0.3), 0.4, 1.3, 0.9),
    ("n16x20x24_s3", (16, 20, 24), 3, (0.5, 0.0, 0.125), 1.0, 2.0, 1.1),
]

TAU_CASES = [
    # name, mesh, nstate, nsup (None: no LSD), omega, tpiba2
    ("n16_s5", (16, 16, 16), 5, None, 1.7, 0.8),
    ("n16x20x24_s4_nsup1", (16, 20, 24), 4, 1, 2.0, 1.1),
]

LSD_CASES = [
    # name, mesh, nstate, nsup, omega, tpiba2
    ("n16_s7_nsup3", (16, 16, 16), 7, 3, 1.3, 0.9),
    ("n20_s6_nsup4", (20, 20, 20), 6, 4, 1.0, 1.0),
]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, nr, ns, fp, omega, tpiba2, (grp, ngrp) in CASES:
        geo = orc.make_geometry(nr)
        c0, f, v = orc.synthetic_inputs(geo, ns, seed=4242 + ns + nr[0], f_pattern=fp)
        rho = orc.rhoofr(geo, c0, f, omega, tpiba2, grp, ngrp)
        c2_in = 0.25 * c0[::-1].copy()
        c2 = orc.vpsi(geo, c0, c2_in, f, v, tpiba2, grp, ngrp)
        np.savez_compressed(os.path.join(out, name + ".npz"), nr=np.array(nr), inyh=geo.inyh, hg=geo.hg,
                            nzhs=geo.nzhs, indzs=geo.indzs, c0=c0, f=f, vpot=v, omega=omega, tpiba2=tpiba2,
                            group=grp, ngroups=ngrp, rhoe=rho["rhoe"], ekin=rho["ekin"],
                            rsum_g=rho["rsum_g"], rsum_r=rho["rsum_r"], c2_in=c2_in, c2_out=c2)
        print(name, "ngw", geo.ngw, "nnr1", geo.nnr1)
    # LSD (cntl%tlsd) fixtures live in tests/golden/lsd/
    os.makedirs(os.path.join(out, "lsd"), exist_ok=True)
    for name, nr, ns, nsup, omega, tpiba2 in LSD_CASES:
        geo = orc.make_geometry(nr)
        c0, f, v = orc.synthetic_inputs(geo, ns, seed=777 + ns + nr[0], f_pattern="mixed")
        v2 = np.stack([v, 0.5 * v[::-1]])
        rho = orc.rhoofr_lsd(geo, c0, f, omega, tpiba2, nsup)
        c2_in = 0.25 * c0[::-1].copy()
        c2 = orc.vpsi_lsd(geo, c0, c2_in, f, v2, tpiba2, nsup)
        np.savez_compressed(os.path.join(out, "lsd", name + ".npz"), nr=np.array(nr), inyh=geo.inyh, hg=geo.hg,
                            c0=c0, f=f, vpot=v2, omega=omega, tpiba2=tpiba2, nsup=nsup, rhoe=rho["rhoe"],
                            ekin=rho["ekin"], rsum_g=rho["rsum_g"], rsum_r=rho["rsum_r"], csums=rho["csums"],
                            csumsabs=rho["csumsabs"], c2_in=c2_in, c2_out=c2)
        print("lsd/" + name, "ngw", geo.ngw)
    # local part of vofrho on the density cutoff (tests/golden/vofrho/): the density comes from rhoofr
    os.makedirs(os.path.join(out, "vofrho"), exist_ok=True)
    for name, nr, omega, tpiba2 in VOFRHO_CASES:
        geo = orc.make_geometry(nr)
        c0, f, _ = orc.synthetic_inputs(geo, 4, seed=99 + nr[0], f_pattern="all2")
        rhoe = orc.rhoofr(geo, c0, f, omega, tpiba2)["rhoe"]
        dgeo = orc.make_density_geometry(nr)
        scg, eivps, eirop = orc.synthetic_vofrho_inputs(dgeo, tpiba2, omega, seed=31 + nr[2])
        r = orc.vofrho_local(dgeo, rhoe, scg, eivps, eirop)
        ener = np.array([r["eh"].real, r["eh"].imag, r["ei"].real, r["ei"].imag, r["ee"].real, r["ee"].imag,
                         r["eps"].real, r["eps"].imag, r["vploc"]])
        np.savez_compressed(os.path.join(out, "vofrho", name + ".npz"), nr=np.array(nr), inyh=dgeo.inyh, hg=dgeo.hg,
                            nzh=dgeo.nzhs, indz=dgeo.indzs, omega=omega, tpiba2=tpiba2, rhoe=rhoe, scg=scg,
                            eivps=eivps, eirop=eirop, rhog=r["rhog"], vtemp=r["vtemp"], v=r["v"], ener=ener)
        print("vofrho/" + name, "nhg", dgeo.ngw)
    # k-points (tests/golden/kpt/): one k-point of rhoofr_c and of vpsi's k-branch
    os.makedirs(os.path.join(out, "kpt"), exist_ok=True)
    for name, nr, ns, kvec, wk, omega, tpiba2 in KPT_CASES:
        geo = orc.make_geometry(nr)
        c0, f, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, ns, kvec=kvec, seed=555 + ns + nr[1])
        rho = orc.rhoofr_kpt(geo, c0, f, wk, hgkp, hgkm, omega, tpiba2)
        c2_in = 0.25 * c0[::-1].copy()
        c2 = orc.vpsi_kpt(geo, c0, c2_in, f, hgkp, hgkm, v, tpiba2)
        np.savez_compressed(os.path.join(out, "kpt", name + ".npz"), nr=np.array(nr), inyh=geo.inyh, hg=geo.hg,
                            c0=c0, f=f, hgkp=hgkp, hgkm=hgkm, wk=wk, vpot=v, omega=omega, tpiba2=tpiba2,
                            rhoe=rho["rhoe"], ekin=rho["ekin"], rsum_g=rho["rsum_g"], c2_in=c2_in, c2_out=c2)
        print("kpt/" + name, "ngw", geo.ngw)
    # meta-GGA (tests/golden/tau/): tauofr + vtaupsi
    os.makedirs(os.path.join(out, "tau"), exist_ok=True)
    for name, nr, ns, nsup, omega, tpiba2 in TAU_CASES:
        geo = orc.make_geometry(nr)
        c0, f, v = orc.synthetic_inputs(geo, ns, seed=321 + ns + nr[2], f_pattern="mixed")
        gk = orc.gk_cartesian(geo)
        tau = orc.tauofr(geo, c0, f, gk, omega, tpiba2, nsup)
        vtau = np.stack([v, 0.5 * v[::-1]])[: (1 if nsup is None else 2)]
        c2_in = 0.25 * c0[::-1].copy()
        c2 = orc.vtaupsi(geo, c0, c2_in, f, gk, vtau, tpiba2, nsup)
        np.savez_compressed(os.path.join(out, "tau", name + ".npz"), nr=np.array(nr), inyh=geo.inyh, hg=geo.hg,
                            c0=c0, f=f, gk=gk, nsup=-1 if nsup is None else nsup, omega=omega, tpiba2=tpiba2,
                            tau=tau, vtau=vtau, c2_in=c2_in, c2_out=c2)
        print("tau/" + name, "ngw", geo.ngw)


if __name__ == "__main__":
    main()
