"""Compare a tuning build of the library against the default build on the same inputs (GPU box).
usage: python tools/gpu_compare.py <other_lib.so> <mesh> <nstate> <pairs_per_batch>"""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
from cpmd_b200 import Plan, lib, synthetic
other = lib.declare(C.CDLL(sys.argv[1]))
n = int(sys.argv[2]); ns = int(sys.argv[3]); mb = int(sys.argv[4])
d = synthetic.make_inputs(n, ns, f_pattern="mixed")
dev = torch.device('cuda:0')
c0 = torch.from_numpy(d['c0']).to(dev); v = torch.from_numpy(d['vpot']).to(dev)
out = []
for cdll in (None, other):
    plan = Plan(d['nr'], d['inyh'], d['hg'], max_batch=mb, _cdll=cdll)
    rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev); c2 = torch.zeros_like(c0)
    e = plan.rhoofr_dev(c0, d['f'], rho); plan.vpsi_dev(c0, c2, d['f'], v); torch.cuda.synchronize()
    out.append((rho.cpu().numpy(), c2.cpu().numpy(), e))
    plan.close()
er = np.abs(out[0][0] - out[1][0]).max() / np.abs(out[0][0]).max()
ec = np.abs(out[0][1] - out[1][1]).max() / np.abs(out[0][1]).max()
print(f'compare {sys.argv[1]}: rho relerr {er:.2e}  c2 relerr {ec:.2e}  scalars {out[0][2]} vs {out[1][2]}')
assert er < 1e-12 and ec < 1e-12
