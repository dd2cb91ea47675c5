"""State-group (CP_GROUPS) decomposition over GPUs: one process per GPU, states block-partitioned
per ``part_1d`` (part_1d.mod.F90:22-57), and the reference's cross-group reductions as
``torch.distributed`` collectives (NCCL over NVLink on GPUs, gloo in CPU tests):

* ``cp_grp_redist(rhoe)``  (rhoofr_utils.mod.F90:457-461, cp_grp_utils.mod.F90:98-120:
  mp_sum over cp_inter_grp)  ->  all_reduce(SUM) of rho(r), FP64;
* the group-partial scalars ekin / rsum_g / rsum_r -> one 3-double all_reduce;
* ``cp_grp_redist(C2_vpsi)`` (vpsi_utils.mod.F90:708-712): the reference sums a zero-padded full
  C2 over groups, i.e. an all-gather of the owned state blocks; offered as :func:`redist_c2`,
  not part of the timed path (north_star keeps C2 sharded by state).

The data path itself has no collective: every GPU holds the full maps, V and a private rho.
"""
from __future__ import annotations

import os


def part_1d_nbr_el_in_blk(n_elem, proc, nproc):
    res = n_elem % nproc
    nbr = (n_elem - res) // nproc
    return nbr + 1 if proc < res else nbr


def part_1d_get_el_in_blk(i_elem, n_elem, proc, nproc):
    res = n_elem % nproc
    nbr = (n_elem - res) // nproc
    return i_elem + nbr * proc + min(proc, res)


def state_block(nstate, group, ngroups):
    """(first 0-based state, count) owned by ``group``."""
    cnt = part_1d_nbr_el_in_blk(nstate, group, ngroups)
    first = part_1d_get_el_in_blk(1, nstate, group, ngroups) - 1 if cnt > 0 else 0
    return first, cnt


def init_from_env(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns
    (rank, world_size, local_rank).  No-op for a single process."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def cp_grp_redist(t, group=None):
    """In-place SUM over the state groups (mp_sum over cp_inter_grp)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def redist_scalars(ekin, rsum_g, rsum_r, device=None, group=None):
    import torch
    import torch.distributed as dist

    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return ekin, rsum_g, rsum_r
    t = torch.tensor([ekin, rsum_g, rsum_r], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    e, g, r = t.tolist()
    return e, g, r


def redist_c2(c2, nstate, group=None):
    """All groups end up with every state's C2 (the reference's sum of zero-padded blocks)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return c2
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    view = torch.view_as_real(c2) if c2.is_complex() else c2
    for g in range(world):
        first, cnt = state_block(nstate, g, world)
        if cnt:
            dist.broadcast(view[first:first + cnt], src=dist.get_global_rank(group, g) if group else g,
                           group=group)
    del me
    return c2


def bcast_potential(v, src=0, group=None):
    """V(r) broadcast once per step (north_star); a no-op when every rank already holds V."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(v, src=src, group=group)
    return v
