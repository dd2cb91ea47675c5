"""cpb_peer_* on real GPUs (needs >= 2 devices on one box: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`; skipped on a single-GPU box): CUDA IPC mapping of the segments,
the all-reduce / broadcast kernels over NVLink against the known answer, bit-identical results on
every rank, and rhoofr on two state groups + cp_grp_redist through the peer kernels == one group."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist

    from cpmd_b200 import Plan, synthetic
    from cpmd_b200.dist import PeerSegment, state_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    ns = 6
    d = synthetic.make_inputs(n, ns, f_pattern="mixed")
    plan = Plan(d["nr"], d["inyh"], d["hg"], device=rank, max_batch=2)
    nn = plan.nnr1 + (plan.nnr1 & 1)
    seg = PeerSegment(2 * nn, rank, world, device=rank)
    rho = seg.tensor(0, plan.nnr1)
    v = seg.tensor(nn, plan.nnr1)
    # known answer first
    rho.copy_(torch.arange(plan.nnr1, dtype=torch.float64, device=dev) * (rank + 1))
    seg.allreduce(0, nn)
    want = torch.arange(plan.nnr1, dtype=torch.float64, device=dev) * (world * (world + 1) / 2)
    assert torch.equal(rho, want)
    if rank == world - 1:
        v.copy_(torch.from_numpy(d["vpot"]).to(dev))
    else:
        v.fill_(-3.0)
    seg.bcast(nn, nn, src=world - 1)
    seg.check()
    assert np.array_equal(v.cpu().numpy(), d["vpot"])
    # the hot path on state groups + the peer all-reduce == one group
    first, cnt = state_block(ns, rank, world)
    c0 = torch.from_numpy(d["c0"]).to(dev)
    sc = plan.rhoofr_dev(c0, d["f"], rho, ngroups=world, my_group=rank)
    for _ in range(2):                                  # twice: epochs advance, buffers are reused
        plan.rhoofr_dev(c0, d["f"], rho, ngroups=world, my_group=rank)
        seg.allreduce(0, nn)
    full = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    plan.rhoofr_dev(c0, d["f"], full)
    err = (rho - full).abs().max().item() / full.abs().max().item()
    assert err < 1e-11, err
    np.save(os.path.join(out_dir, f"rho_{rank}.npy"), rho.cpu().numpy())
    seg.barrier()
    dist.barrier()
    del rho, v
    seg.close()
    dist.destroy_process_group()
    del sc, first, cnt


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_collectives_over_nvlink(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    import torch.multiprocessing as mp

    port = 29600 + world
    mp.spawn(_worker, args=(world, port, 48, str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "rho_0.npy")
    for r in range(1, world):
        assert np.array_equal(np.load(tmp_path / f"rho_{r}.npy"), ref)     # bit-identical on every rank
