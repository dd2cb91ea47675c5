"""Host-pointer (Fortran drop-in) path timing: PCIe copy rates and the two host entry points.
usage: python tools/e2e_probe.py <mesh> <nstate> <pairs_per_batch>"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from cpmd_b200 import Plan, lib, synthetic
n = int(sys.argv[1]); ns = int(sys.argv[2]); mb = int(sys.argv[3])
d = synthetic.make_inputs(n, ns)
plan = Plan(d['nr'], d['inyh'], d['hg'], max_batch=mb)
dev = torch.device('cuda:0')
c0h = torch.from_numpy(d['c0']).pin_memory(); c2h = torch.zeros_like(c0h).pin_memory()
vh = torch.from_numpy(d['vpot']).pin_memory(); rhoh = torch.empty(plan.nnr1, dtype=torch.float64).pin_memory()
c0d = torch.empty_like(c0h, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); a = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time() - a) / reps
gb = c0h.numel() * 16 / 1e9
print(f'c0 block {gb:.2f} GB; H2D {gb / t(lambda: c0d.copy_(c0h, non_blocking=True)):.1f} GB/s; '
      f'D2H {gb / t(lambda: c2h.copy_(c0d, non_blocking=True)):.1f} GB/s', flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): c0d.copy_(c0h, non_blocking=True)
    with torch.cuda.stream(s2): c2h.copy_(c0d, non_blocking=True)
print(f'bidirectional: {2 * gb / t(both):.1f} GB/s aggregate', flush=True)
f = d['f']
for name, fn in (('rhoofr (upload c0, KEEP)', lambda: plan.rhoofr(c0h, f, rhoh, flags=lib.CPB_C0_KEEP)),
                 ('vpsi += (REUSE)', lambda: plan.vpsi(c0h, c2h, f, vh, flags=lib.CPB_C0_REUSE)),
                 ('vpsi overwrite (REUSE)', lambda: plan.vpsi(c0h, c2h, f, vh, flags=lib.CPB_C0_REUSE | lib.CPB_VPSI_OVERWRITE)),
                 ('rhoofr (REUSE)', lambda: plan.rhoofr(c0h, f, rhoh, flags=lib.CPB_C0_REUSE | lib.CPB_C0_KEEP))):
    print(f'{name}: {t(fn) * 1e3:.1f} ms', flush=True)
