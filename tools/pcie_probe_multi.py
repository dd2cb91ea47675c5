"""Concurrent host<->device copy bandwidth of all ranks of one box (why does the e2e
number not scale with the GPU count?).  Launch with torchrun; every rank copies pinned buffers of the size of
its c0 block up and down at the same time as the others, alone and together, and prints its rates; rank 0
also prints the CPU affinity mask and `nvidia-smi topo -m`.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe_multi.py [MB]
"""
import os
import subprocess
import sys

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 237
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
dev = torch.device("cuda", local)
n = mb * (1 << 20) // 8
hin = torch.empty(n, dtype=torch.float64).pin_memory()
hout = torch.empty(n, dtype=torch.float64).pin_memory()
hin.fill_(1.0)
din = torch.empty(n, dtype=torch.float64, device=dev)
dout = torch.ones(n, dtype=torch.float64, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(up, down, reps=8, solo_rank=None):
    """GB/s of this rank; solo_rank: only that rank copies (the others idle)."""
    active = solo_rank is None or solo_rank == rank
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if active:
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_in):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s_out):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_out)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    gb = reps * (int(up) + int(down)) * n * 8 / 1e9
    return gb / (ms * 1e-3) if active else 0.0


def gather(x):
    if world == 1:
        return [x]
    out = [None] * world
    dist.all_gather_object(out, x)
    return out


if rank == 0:
    try:
        print("affinity:", sorted(os.sched_getaffinity(0)))
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    except Exception as e:  # noqa: BLE001
        print("topology query failed:", e)
for name, up, down in (("H2D", True, False), ("D2H", False, True), ("duplex", True, True)):
    run(up, down, reps=2)
    solo = [run(up, down, solo_rank=r) for r in range(world)]
    solo = [max(gather(s)) for s in solo]
    together = gather(run(up, down))
    if rank == 0:
        print(f"{name:7s} {mb} MB buffers: alone per rank " + " ".join(f"{g:5.1f}" for g in solo) +
              f" GB/s | all {world} ranks at once " + " ".join(f"{g:5.1f}" for g in together) +
              f" GB/s, aggregate {sum(together):6.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
