"""Golden vectors made by the REFERENCE'S OWN STATEMENTS (oracle/fsnip.py executes the Fortran line ranges of
rhoofr_utils / vpsi_utils / density_utils / kin_energy_utils / dotp_utils / part_1d from /root/reference/src; the
transforms between them come from the oracle).  Run here, where the reference tree exists; the fixtures travel.
usage: python tools/make_golden_fsnip.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import cpmd_oracle as orc          # noqa: E402
from oracle import fsnip, fsnip_cases as fc    # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fsnip")

CASES = [
    # name, nr, nstate, f_pattern, gcutw scale, b, omega, tpiba2, ngroups, group, tksham
    ("n16_mixed", (16, 16, 16), 5, "mixed", 1.0, None, 2.5, 1.3, 1, 0, False),
    ("n16_mixed_tksham", (16, 16, 16), 5, "mixed", 1.0, None, 2.5, 1.3, 1, 0, True),
    ("n16_group1of3", (16, 16, 16), 7, "mixed", 1.0, None, 1.0, 1.0, 3, 1, False),
    ("aniso_16_20_24", (16, 20, 24), 4, "all2", 1.0, None, 3.1, 0.7, 1, 0, False),
    ("cell_n20", (20, 20, 20), 5, "mixed", 0.6, [[1.0, 0.0, 0.0], [0.27, 1.06, 0.0], [0.14, -0.21, 0.93]], 41.7, 0.83,
     2, 0, False),
]


def main():
    assert fsnip.available(), "needs /root/reference/src"
    os.makedirs(OUT, exist_ok=True)
    for name, nr, ns, fp, scale, b, omega, tpiba2, ngroups, group, tksham in CASES:
        geo = orc.make_geometry(nr, gcutw=scale * (min(nr) / 4.0) ** 2, b=None if b is None else np.array(b))
        c0, f, v = orc.synthetic_inputs(geo, ns, seed=len(name) * 7 + ns, f_pattern=fp)
        ekin, rsum = fc.kin_energy(geo, c0, f, tpiba2)
        rho = fc.rhoofr(geo, c0, f, omega, tpiba2, group, ngroups)
        c2_in = 0.5 * c0
        c2 = fc.vpsi(geo, c0, c2_in, f, v, tpiba2, group, ngroups, tksham)
        pairs_r = np.array(fc.pair_loop("rhoofr", ns, group, ngroups), dtype=np.int32).reshape(-1, 2)
        pairs_v = np.array(fc.pair_loop("vpsi", ns, group, ngroups), dtype=np.int32).reshape(-1, 2)
        rsum_r = rho.sum() * omega / float(np.prod(nr))                  # rhoofr_utils.mod.F90:612-617
        np.savez_compressed(os.path.join(OUT, name + ".npz"), nr=np.array(nr), inyh=geo.inyh, hg=geo.hg, c0=c0, f=f,
                            vpot=v, omega=omega, tpiba2=tpiba2, ngroups=ngroups, group=group, tksham=tksham,
                            ekin=ekin, rsum_g=rsum, rsum_r=rsum_r, rhoe=rho, c2_in=c2_in, c2_out=c2,
                            pairs_rhoofr=pairs_r, pairs_vpsi=pairs_v)
        print(name, "ngw", geo.ngw, "ekin", ekin, "rsum", rsum)


if __name__ == "__main__":
    main()
