"""cpb_peer_* on real GPUs: CUDA IPC mapping of the segments, the all-reduce / broadcast / all-gather
kernels against the known answer, bit-identical results on every rank, rhoofr on state groups +
cp_grp_redist through the peer kernels == one group, vpsi with the V broadcast overlapped
(cpb_plan_set_vpot_event) and cp_grp_redist(C2) == one group.

Two forms: one process per GPU over NVLink (needs >= 2 devices on one box: `gpurun --gpus 2 -- python -m
pytest tests/test_gpu_multi.py -m gpu`; skipped otherwise), and two processes sharing ONE GPU (the same
IPC mapping, flag barriers and kernels; the segments then are peer memory of the same device), which
runs on a single-GPU box too.  Logs of the multi-GPU runs are kept under profiles/."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _worker(rank, world, port, n, out_dir, devs):
    import torch.distributed as dist

    from cpmd_b200 import Plan, synthetic
    from cpmd_b200.dist import PeerSegment, state_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    di = devs[rank]
    torch.cuda.set_device(di)
    dev = torch.device("cuda", di)
    ns = 6
    d = synthetic.make_inputs(n, ns, f_pattern="mixed")
    plan = Plan(d["nr"], d["inyh"], d["hg"], device=di, max_batch=2)
    ngw = plan.ngw
    nn = plan.nnr1 + (plan.nnr1 & 1)
    seg = PeerSegment(2 * nn + 2 * ns * ngw, rank, world, device=di)
    seg.set_timeout_ms(120000.0)                        # ranks sharing one GPU are time-sliced
    rho = seg.tensor(0, plan.nnr1)
    v = seg.tensor(nn, plan.nnr1)
    c2 = torch.view_as_complex(seg.tensor(2 * nn, 2 * ns * ngw).view(ns, ngw, 2))
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
    for _ in range(2):                                  # twice: epochs advance, buffers are reused
        ek, rg, rr = plan.rhoofr_dev(c0, d["f"], rho, ngroups=world, my_group=rank)
        seg.allreduce(0, nn)
    ek, rg, rr = seg.allreduce_scalars([ek, rg, rr])    # group-partial scalars, rank order
    full = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    ek1, rg1, rr1 = plan.rhoofr_dev(c0, d["f"], full)
    err = (rho - full).abs().max().item() / full.abs().max().item()
    assert err < 1e-11, err
    assert abs(ek - ek1) < 1e-9 * max(1.0, abs(ek1)) and abs(rg - rg1) < 1e-9 and abs(rr - rr1) < 1e-9
    # vpsi on the group's block with the V broadcast on a side stream (only the z pass waits for it),
    # then cp_grp_redist(C2) (vpsi_utils.mod.F90:708-712) == the one-group result (the pairing of states
    # differs between the groupings, so to rounding; the ranks among themselves agree bit for bit)
    side = torch.cuda.Stream(device=dev)
    if rank != 0:
        v.fill_(-5.0)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.stream(side):
        seg.bcast(nn, nn, src=0, stream=side)
        ev = torch.cuda.Event()
        ev.record(side)
    c2.zero_()
    plan.set_vpot_event(ev)
    plan.vpsi_dev(c0, c2, d["f"], v, ngroups=world, my_group=rank)
    seg.redist_c2(2 * nn, ngw, ns)
    seg.check()
    c2_one = torch.zeros_like(c0)
    plan.vpsi_dev(c0, c2_one, d["f"], v)
    torch.cuda.synchronize()
    err = (c2 - c2_one).abs().max().item() / c2_one.abs().max().item()
    assert err < 1e-11, err
    np.save(os.path.join(out_dir, f"rho_{rank}.npy"), rho.cpu().numpy())
    np.save(os.path.join(out_dir, f"c2_{rank}.npy"), c2.cpu().numpy())
    seg.barrier()
    dist.barrier()
    del rho, v, c2
    seg.close()
    dist.destroy_process_group()
    del first, cnt


def _check_ranks_identical(tmp_path, world):
    for what in ("rho", "c2"):
        ref = np.load(tmp_path / f"{what}_0.npy")
        for r in range(1, world):
            assert np.array_equal(np.load(tmp_path / f"{what}_{r}.npy"), ref)     # bit-identical on every rank


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_collectives_over_nvlink(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    import torch.multiprocessing as mp

    port = 29600 + world
    mp.spawn(_worker, args=(world, port, 48, str(tmp_path), list(range(world))), nprocs=world, join=True)
    _check_ranks_identical(tmp_path, world)


def test_peer_collectives_two_processes_one_gpu(tmp_path):
    """The same worker with both ranks on device 0: runs on the driver's single-GPU box."""
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, 29611, 36, str(tmp_path), [0, 0]), nprocs=2, join=True)
    _check_ranks_identical(tmp_path, 2)


def _late_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from cpmd_b200.dist import PeerSegment

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    seg = PeerSegment(64, rank, world, device=0)
    seg.set_timeout_ms(300.0)
    a = seg.tensor(0, 64)
    a.fill_(float(rank + 1))
    torch.cuda.synchronize()
    dist.barrier()
    msg = "ok"
    if rank == 0:
        seg.allreduce(0, 64)                            # rank 1 never shows up
        try:
            seg.check()
        except RuntimeError as e:
            msg = str(e)
        assert torch.all(a == 1.0)                      # untouched, not half-summed
    dist.barrier()
    if rank == 1:
        try:
            seg.check()                                 # the waiting rank raised this rank's error word too
        except RuntimeError as e:
            msg = str(e)
    open(os.path.join(out_dir, f"msg_{rank}.txt"), "w").write(msg)
    dist.barrier()
    del a
    seg.close()
    dist.destroy_process_group()


def test_peer_barrier_timeout_is_reported(tmp_path):
    """a missing rank must surface as an error on every rank, with the data left alone."""
    import torch.multiprocessing as mp

    mp.spawn(_late_worker, args=(2, 29613, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert "timeout" in open(tmp_path / f"msg_{r}.txt").read()


def test_one_process_two_devices():
    """plans on two devices in ONE process - the opt-in to > 48 KB of dynamic shared memory
    is per device (the 192-point kernels need it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one box")
    from cpmd_b200 import Plan, synthetic

    d = synthetic.make_inputs(192, 4)
    out = []
    for di in (0, 1):
        dev = torch.device("cuda", di)
        plan = Plan(d["nr"], d["inyh"], d["hg"], device=di, max_batch=2)
        with torch.cuda.device(dev):
            c0 = torch.from_numpy(d["c0"]).to(dev)
            rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
            plan.rhoofr_dev(c0, d["f"], rho)
            c2 = torch.zeros_like(c0)
            plan.vpsi_dev(c0, c2, d["f"], torch.from_numpy(d["vpot"]).to(dev))
            torch.cuda.synchronize(dev)
        out.append((rho.cpu(), c2.cpu()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
