"""World-size-2 (and 3) gloo runs of the state-group path on CPU: every rank runs its block of
states through the kernel simulator build, then cp_grp_redist (all_reduce) combines rho and the
scalars exactly like cpmd_b200.dist does over NCCL on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nstate, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes

    from cpmd_b200 import dist as cdist
    from cpmd_b200 import lib, synthetic
    from cpmd_b200.api import Plan

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = cdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    cdll = lib.declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libcpb200_emu.so")))
    d = synthetic.make_inputs(n, nstate, f_pattern="mixed")
    plan = Plan(d["nr"], d["inyh"], d["hg"], max_batch=2, _cdll=cdll)
    rho, ekin, rg, rr = plan.rhoofr(d["c0"], d["f"], ngroups=world, my_group=rank)
    rho_t = torch.from_numpy(rho)
    cdist.cp_grp_redist(rho_t)
    ekin, rg, rr = cdist.redist_scalars(ekin, rg, rr)
    v = torch.from_numpy(d["vpot"].copy())
    if rank != 0:
        v.zero_()
    cdist.bcast_potential(v, src=0)
    c2 = np.zeros_like(d["c0"])
    plan.vpsi(d["c0"], c2, d["f"], v.numpy(), ngroups=world, my_group=rank)
    c2_t = torch.from_numpy(c2)
    first, cnt = cdist.state_block(nstate, rank, world)
    assert not c2[:first].any() and not c2[first + cnt:].any()      # C2 stays sharded by state
    cdist.redist_c2(c2_t, nstate)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rho=rho_t.numpy(), c2=c2_t.numpy(),
             s=np.array([ekin, rg, rr]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nstate", [(2, 6), (3, 7)])
def test_state_groups_over_gloo(emu_cdll, tmp_path, world, nstate):
    from helpers import ETOL, RTOL, relmax
    from cpmd_b200 import synthetic
    from oracle import cpmd_oracle as orc

    n = 16
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, nstate, str(tmp_path)), nprocs=world, join=True)
    d = synthetic.make_inputs(n, nstate, f_pattern="mixed")
    geo = orc.make_geometry(n)
    ref = orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)
    c2_ref = orc.vpsi(geo, d["c0"], np.zeros_like(d["c0"]), d["f"], d["vpot"], 1.0)
    outs = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for o in outs:
        assert relmax(o["rho"], ref["rhoe"]) < RTOL
        assert relmax(o["c2"], c2_ref) < RTOL
        assert np.abs(o["s"] - (ref["ekin"], ref["rsum_g"], ref["rsum_r"])).max() < ETOL
    # every rank holds bit-identical reduced results
    for o in outs[1:]:
        assert np.array_equal(o["rho"], outs[0]["rho"])
