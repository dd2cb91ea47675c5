import sys, time, json
sys.path.insert(0,'.')
import numpy as np, torch
from cpmd_b200 import Plan, synthetic
n=int(sys.argv[1]); ns=int(sys.argv[2]); mb=int(sys.argv[3])
t0=time.time(); d=synthetic.make_inputs(n,ns); print('gen',time.time()-t0, flush=True)
plan=Plan(d['nr'],d['inyh'],d['hg'],max_batch=mb); print(plan.info, flush=True)
dev=torch.device('cuda:0')
c0=torch.from_numpy(d['c0']).to(dev); v=torch.from_numpy(d['vpot']).to(dev)
rho=torch.empty(plan.nnr1,dtype=torch.float64,device=dev); c2=torch.zeros_like(c0)
for it in range(3):
    torch.cuda.synchronize(); t=time.time(); e=plan.rhoofr_dev(c0,d['f'],rho); torch.cuda.synchronize(); t1=time.time()-t
    t=time.time(); plan.vpsi_dev(c0,c2,d['f'],v); torch.cuda.synchronize(); t2=time.time()-t
    print(f'it{it} rhoofr {t1*1e3:.2f} ms vpsi {t2*1e3:.2f} ms  step {(t1+t2)*1e3:.2f} ms  bandFFT/s {3*ns/(t1+t2):.0f}', e, flush=True)
