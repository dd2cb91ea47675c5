"""Quick device-resident timing of rhoofr + vpsi with the per-kernel-class breakdown.
usage: python tools/gpu_probe.py <mesh> <nstate> <pairs_per_batch> [iters]
env CPB_STREAMS=1|2 selects serialised / overlapped batches; the per-kernel breakdown is always
taken in one extra serialised iteration."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from cpmd_b200 import Plan, synthetic
n = int(sys.argv[1]); ns = int(sys.argv[2]); mb = int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = synthetic.make_inputs(n, ns)
plan = Plan(d['nr'], d['inyh'], d['hg'], max_batch=mb)
print(plan.info, flush=True)
dev = torch.device('cuda:0')
c0 = torch.from_numpy(d['c0']).to(dev); v = torch.from_numpy(d['vpot']).to(dev)
rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev); c2 = torch.zeros_like(c0)
def one(tag):
    torch.cuda.synchronize(); t = time.time(); e = plan.rhoofr_dev(c0, d['f'], rho); torch.cuda.synchronize(); t1 = time.time() - t
    t = time.time(); plan.vpsi_dev(c0, c2, d['f'], v); torch.cuda.synchronize(); t2 = time.time() - t
    print(f'{tag} rhoofr {t1*1e3:.2f} ms vpsi {t2*1e3:.2f} ms  step {(t1+t2)*1e3:.2f} ms  bandFFT/s {3*ns/(t1+t2):.0f}', e, flush=True)
for it in range(iters):
    one(f'it{it}')
plan.set_streams(1); plan.set_profiling(True)
one('serialised+profiled')
kt = plan.kernel_times(reset=True)
print('kernel ms (serialised iteration): ' + '  '.join(f'{k} {v_[0]:.2f}/{v_[1]}' for k, v_ in kt.items()), flush=True)
