"""One small rhoofr + vpsi (+ LSD, k-point, vofrho_local) run on cuda:0, checked against the oracle; meant to be
run under compute-sanitizer (tools/run_sanitizer.sh): memcheck, racecheck, synccheck, initcheck.
usage: python tools/sanitize_case.py <mesh> <nstate> [full]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from cpmd_b200 import Plan, synthetic
from oracle import cpmd_oracle as orc

n = int(sys.argv[1]); ns = int(sys.argv[2]); full = len(sys.argv) > 3
d = synthetic.make_inputs(n, ns, f_pattern="mixed")
geo = orc.make_geometry(n)
plan = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], device=0, max_batch=2)
dev = torch.device("cuda:0")
c0 = torch.from_numpy(d["c0"]).to(dev); v = torch.from_numpy(d["vpot"]).to(dev)
rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
ekin, rg, rr = plan.rhoofr_dev(c0, d["f"], rho)
c2 = torch.from_numpy(0.5 * d["c0"]).to(dev)   # by memcpy: initcheck only sees the library's kernels and the copies
plan.vpsi_dev(c0, c2, d["f"], v)
torch.cuda.synchronize()
ref = orc.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"])
c2ref = orc.vpsi(geo, d["c0"], 0.5 * d["c0"], d["f"], d["vpot"], d["tpiba2"])
e1 = np.abs(rho.cpu().numpy() - ref["rhoe"]).max() / np.abs(ref["rhoe"]).max()
e2 = np.abs(c2.cpu().numpy() - c2ref).max() / np.abs(c2ref).max()
assert e1 < 1e-11 and e2 < 1e-11 and abs(ekin - ref["ekin"]) < 1e-9, (e1, e2)
msg = f"mesh {n} states {ns}: rho {e1:.1e} c2 {e2:.1e}"
if full:
    # host-pointer entry points (staging copies on their own streams) and the LSD variant
    rho_h, ek_h, _, _ = plan.rhoofr(d["c0"], d["f"])
    c2_h = 0.5 * d["c0"]
    plan.vpsi(d["c0"], c2_h, d["f"], d["vpot"])
    assert np.array_equal(rho_h, rho.cpu().numpy()) and np.array_equal(c2_h, c2.cpu().numpy())
    nsup = ns // 2 + 1
    refl = orc.rhoofr_lsd(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], nsup)
    out = plan.rhoofr_lsd(d["c0"], d["f"], nsup)
    e3 = np.abs(out[0] - refl["rhoe"]).max() / np.abs(refl["rhoe"]).max()
    assert e3 < 1e-11, e3
    msg += f" host forms bit-identical, lsd {e3:.1e}"
print("sanitize case ok:", msg)
