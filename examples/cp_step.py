#!/usr/bin/env python3
"""One electronic step of the hot path on one GPU, device resident, through the public API:

    rhoofr (density)  ->  vofrho_local (Hartree + local pseudopotential, G space)  ->  vpsi (V psi)

with synthetic plane-wave coefficients (no CPMD input files are needed).  The exchange-correlation
part of vofrho is outside the library: a real driver adds v_xc(r) to V before calling vpsi.

    python examples/cp_step.py [mesh] [states]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpmd_b200 import Plan, gvec, synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    nstate = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda:0")
    d = synthetic.make_inputs(n, nstate)                       # inyh, hg, c0, f, ... like the cppt module
    wplan = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"])            # wavefunction cutoff
    inyh_d, hg_d = gvec.half_sphere(d["nr"], (n / 2.0) ** 2)                      # density cutoff (dual 4)
    dplan = Plan(d["nr"], inyh_d, hg_d, d["tpiba2"], d["omega"], max_batch=1)
    print(f"mesh {n}^3: ngw {wplan.ngw}, nhg {dplan.ngw}, {nstate} states, "
          f"work space {wplan.info['workspace_bytes'] / 1e6:.0f} MB")

    c0 = torch.from_numpy(d["c0"]).to(dev)
    c2 = torch.zeros_like(c0)
    rho = torch.empty(wplan.nnr1, dtype=torch.float64, device=dev)
    # G-space inputs of ppener: Coulomb kernel, and (here) no ionic terms
    scg = torch.zeros(dplan.ngw, dtype=torch.float64)
    scg[1:] = 4.0 * np.pi / (d["tpiba2"] * torch.from_numpy(hg_d[1:]))
    scg = scg.to(dev)
    zero = torch.zeros(dplan.ngw, dtype=torch.complex128, device=dev)

    ekin, csumg, csumr = wplan.rhoofr_dev(c0, d["f"], rho)                        # rho(r), E_kin, charge
    e = dplan.vofrho_local_dev(rho, scg, zero, zero, rho)                         # rho(r) -> V_H(r), in place
    wplan.vpsi_dev(c0, c2, d["f"], rho)                                           # c2 += -f (T + V) c0
    torch.cuda.synchronize()
    eh = e["eh"].real * d["omega"]
    print(f"charge {csumg:.10f} (G) / {csumr:.10f} (r)   E_kin {ekin:.10f}   E_Hartree {eh:.10f}")
    w = torch.full((wplan.ngw,), 2.0, dtype=torch.float64, device=dev)
    w[0] = 1.0
    band = -(w * (c0.real * c2.real + c0.imag * c2.imag)).sum().item()
    print(f"-sum_i <c0_i|c2_i> = {band:.10f}  (= E_kin + 2 E_Hartree = {ekin + 2 * eh:.10f})")


if __name__ == "__main__":
    main()
