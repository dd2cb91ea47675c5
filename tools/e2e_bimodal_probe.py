"""Is the duplex phase of cpb_vpsi (+=) bimodal per plan (its streams) or per host buffer?  Re-creates the plan /
the pinned buffers in one process and times the call."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from cpmd_b200 import Plan, lib, synthetic
d = synthetic.make_inputs(192, 512)
f = d['f']
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); a = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time() - a) / reps * 1e3
c0h = torch.from_numpy(d['c0']).pin_memory(); vh = torch.from_numpy(d['vpot']).pin_memory()
for trial in range(6):
    plan = Plan(d['nr'], d['inyh'], d['hg'], max_batch=32)
    rhoh = torch.empty(plan.nnr1, dtype=torch.float64).pin_memory()
    plan.rhoofr(c0h, f, rhoh, flags=lib.CPB_C0_KEEP)
    out = []
    for k in range(3):
        c2h = torch.zeros_like(c0h).pin_memory()
        out.append(t(lambda: plan.vpsi(c0h, c2h, f, vh, flags=lib.CPB_C0_REUSE)))
        del c2h
    print(f"plan {trial}: vpsi += with three fresh c2 buffers: " + " ".join(f"{x:.1f}" for x in out) + " ms", flush=True)
    del plan
