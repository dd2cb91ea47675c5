"""Multi-GPU step breakdown (torchrun): device time of rhoofr, the rho allreduce, the V broadcast
and vpsi, each bracketed by CUDA events, max over ranks.  usage (under torchrun):
  python -m torch.distributed.run --nproc-per-node N tools/scale_probe.py [mesh] [states] [batch]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from cpmd_b200 import Plan, dist as cdist, synthetic
rank, world, local = cdist.init_from_env()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 512
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
d = synthetic.make_inputs(n, ns)
first, cnt = cdist.state_block(ns, rank, world)
plan = Plan(d['nr'], d['inyh'], d['hg'], device=local, max_batch=mb)
c0 = torch.from_numpy(d['c0'][first:first + cnt]).to(dev); c2 = torch.zeros_like(c0)
f = np.ascontiguousarray(d['f'][first:first + cnt])
mode = os.environ.get('CPB_COLLECTIVES', 'peer') if world > 1 else 'none'
nn = plan.nnr1 + (plan.nnr1 & 1)
if mode == 'peer':
    seg = cdist.PeerSegment(2 * nn, rank, world, device=local)
    rho = seg.tensor(0, plan.nnr1); v = seg.tensor(nn, plan.nnr1); v.copy_(torch.from_numpy(d['vpot']).to(dev))
else:
    v = torch.from_numpy(d['vpot']).to(dev); rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream()
names = ['rhoofr', 'allreduce', 'bcast', 'vpsi', 'step']
acc = {k: 0.0 for k in names}
def ev(): return torch.cuda.Event(enable_timing=True)
iters = 8
for it in range(3 + iters):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e = [ev() for _ in range(5)]
    e[0].record(); plan.rhoofr_dev(c0, f, rho, stream=st)
    e[1].record(); seg.allreduce(0, nn, stream=st) if mode == 'peer' else cdist.cp_grp_redist(rho)
    e[2].record(); seg.bcast(nn, nn, src=0, stream=st) if mode == 'peer' else cdist.bcast_potential(v, src=0)
    e[3].record(); plan.vpsi_dev(c0, c2, f, v, stream=st)
    e[4].record(); torch.cuda.synchronize()
    if it >= 3:
        for i, k in enumerate(names[:4]): acc[k] += e[i].elapsed_time(e[i + 1])
        acc['step'] += e[0].elapsed_time(e[4])
t = torch.tensor([acc[k] / iters for k in names], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f'collectives={mode} N={world} mesh {n} states {ns}: ' + '  '.join(f'{k} {x:.3f} ms' for k, x in zip(names, t.tolist())), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
