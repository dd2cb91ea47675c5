"""Collective micro-benchmark (torchrun): cpb_peer_* kernels vs torch.distributed (NCCL) on the rho /
V arrays of the given mesh.  Device time between CUDA events, max over ranks, median of the iterations.
  python -m torch.distributed.run --nproc-per-node N tools/peer_probe.py [mesh]"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from cpmd_b200 import dist as cdist
rank, world, local = cdist.init_from_env()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
kr = n + 1
nn = kr ** 3 + (kr ** 3 & 1)
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
seg = cdist.PeerSegment(2 * nn, rank, world, device=local)
a = seg.tensor(0, nn); b = torch.empty(nn, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream()
def run(fn, iters=20):
    ts, ws = [], []
    for it in range(iters + 3):
        a.fill_(rank + 1.0); b.fill_(rank + 1.0)
        dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
        if it >= 3: ts.append(e0.elapsed_time(e1)); ws.append((w1 - w0) * 1e3)
    t = torch.tensor([np.median(ts), np.median(ws)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
res = {}
res['peer allreduce'] = run(lambda: seg.allreduce(0, nn, stream=st))
assert abs(a[5].item() - world * (world + 1) / 2) < 1e-12
res['nccl allreduce'] = run(lambda: dist.all_reduce(b))
res['peer bcast'] = run(lambda: seg.bcast(0, nn, src=0, stream=st))
assert a[7].item() == 1.0
res['nccl bcast'] = run(lambda: dist.broadcast(b, src=0))
res['peer barrier'] = run(lambda: seg.barrier(stream=st))
if rank == 0:
    print(f'N={world} {nn * 8 / 1e6:.1f} MB: ' + '  '.join(f'{k} {v[0]:.3f} ms (wall {v[1]:.3f})' for k, v in res.items()), flush=True)
dist.barrier(); del a; seg.close(); dist.destroy_process_group()
