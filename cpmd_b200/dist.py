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


# ---------------------------------------------------------------------------------------------
# cp_grp_redist / V broadcast as hand-written kernels over NVLink peer memory (cpb_peer_*)
# ---------------------------------------------------------------------------------------------

class _DevArray:
    """__cuda_array_interface__ view of raw device memory (float64, 1-D) for torch.as_tensor."""

    def __init__(self, ptr, n, owner):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerSegment:
    """One rank's segment of NVLink-mapped device memory plus the collectives that run on it
    (``include/cpb200.h``: cpb_peer_*).  Arrays that take part in a collective are carved out of the
    segment with :meth:`tensor`; offsets and sizes are in doubles.

    ``exchange`` all-gathers the 64-byte handles: by default ``torch.distributed.all_gather_object``
    over ``group``; tests pass their own (threads of one process on the kernel simulator)."""

    def __init__(self, ndoubles, rank, world, device=0, exchange=None, group=None, _cdll=None):
        import ctypes as C

        from . import lib as _lib

        self._L = _cdll if _cdll is not None else _lib.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        self.n = int(ndoubles) + (int(ndoubles) & 1)
        self._h = C.c_void_p()
        handle = (C.c_ubyte * _lib.CPB_PEER_HANDLE_BYTES)()
        # Every rank takes part in both exchanges even if its own step failed, so that a failure
        # raises on ALL ranks (the caller may then choose another collective path) instead of
        # leaving the others blocked in the exchange.
        rc = self._L.cpb_peer_create(C.byref(self._h), self.device, self.rank, self.world, self.n * 8, handle)
        err = None if rc == 0 else f"cpb_peer_create: {self._L.cpb_peer_last_error().decode()}"
        mine = bytes(handle) if rc == 0 else None
        if exchange is None:
            import torch.distributed as dist

            def exchange(b):
                out = [None] * self.world
                dist.all_gather_object(out, b, group=group)
                return out
        blobs = exchange(mine) if self.world > 1 else [mine]
        if any(b is None for b in blobs):
            self.close()
            raise RuntimeError(err or "cpb_peer_create failed on another rank")
        allh = b"".join(blobs)
        assert len(allh) == self.world * _lib.CPB_PEER_HANDLE_BYTES
        buf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
        rc = self._L.cpb_peer_connect(self._h, buf)
        err = None if rc == 0 else f"cpb_peer_connect: {self._L.cpb_peer_last_error().decode()}"
        oks = exchange(b"ok" if rc == 0 else None) if self.world > 1 else [b"ok" if rc == 0 else None]
        if any(b is None for b in oks):
            self.close()
            raise RuntimeError(err or "cpb_peer_connect failed on another rank")
        self.ptr = int(self._L.cpb_peer_local_ptr(self._h))
        self._views = []          # weak references to the tensors handed out by tensor()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"cpb_peer error {rc}: {self._L.cpb_peer_last_error().decode()}")

    def tensor(self, offset, n):
        """float64 CUDA tensor over doubles [offset, offset+n) of the local segment."""
        import torch

        import weakref

        if offset < 0 or offset + n > self.n:
            raise ValueError("range outside the segment")
        t = torch.as_tensor(_DevArray(self.ptr + 8 * offset, n, self), device=torch.device("cuda", self.device))
        self._views.append(weakref.ref(t))
        return t

    def numpy(self, offset, n):
        """Host view (kernel simulator only: its "device" memory is host memory)."""
        import ctypes as C

        import numpy as np
        return np.ctypeslib.as_array((C.c_double * n).from_address(self.ptr + 8 * offset))

    @staticmethod
    def _sp(stream):
        import ctypes as C
        if stream is None:
            return None
        return C.c_void_p(stream if isinstance(stream, int) else stream.cuda_stream)

    def allreduce(self, offset, n, stream=None):
        """In-place sum over the ranks = cp_grp_redist (rhoofr_utils.mod.F90:457-461).  Enqueues only."""
        self._check(self._L.cpb_peer_allreduce_f64(self._h, int(offset), int(n), self._sp(stream)))

    def bcast(self, offset, n, src=0, stream=None):
        self._check(self._L.cpb_peer_bcast_f64(self._h, int(offset), int(n), int(src), self._sp(stream)))

    def allgather(self, offset, counts, stream=None):
        """In place: block q (counts[q] doubles, back to back from ``offset``) of rank q to every rank."""
        import ctypes as C
        if len(counts) != self.world:
            raise ValueError("one count per rank")
        arr = (C.c_size_t * self.world)(*[int(c) for c in counts])
        self._check(self._L.cpb_peer_allgather_f64(self._h, int(offset), arr, self._sp(stream)))

    def redist_c2(self, offset, ld, nstate, stream=None):
        """cp_grp_redist(C2_vpsi) (vpsi_utils.mod.F90:708-712): all-gather of the part_1d state blocks of
        the (nstate, ld) complex128 array that starts ``offset`` doubles into the segment."""
        self._check(self._L.cpb_peer_redist_c2(self._h, int(offset), int(ld), int(nstate), self._sp(stream)))

    def allreduce_scalars(self, vals, stream=None):
        """Sum of up to 8 host doubles over the ranks, in rank order (ekin, rsum_g, rsum_r)."""
        import ctypes as C
        vals = [float(v) for v in vals]
        arr = (C.c_double * len(vals))(*vals)
        self._check(self._L.cpb_peer_allreduce_scalars(self._h, arr, len(vals), self._sp(stream)))
        return list(arr)

    def set_timeout_ms(self, ms):
        self._check(self._L.cpb_peer_set_timeout_ms(self._h, float(ms)))

    def barrier(self, stream=None):
        self._check(self._L.cpb_peer_barrier(self._h, self._sp(stream)))

    def check(self, stream=None):
        """Synchronise the stream and raise if a barrier of an earlier collective timed out."""
        self._check(self._L.cpb_peer_check(self._h, self._sp(stream)))

    def close(self, force=False):
        """Unmap the peers and free the own segment.  Refuses while tensors handed out by :meth:`tensor`
        are still referenced (they would dangle) unless ``force``."""
        if getattr(self, "_h", None) is not None and self._h.value:
            alive = [r for r in getattr(self, "_views", []) if r() is not None]
            if alive and not force:
                raise RuntimeError(f"PeerSegment.close(): {len(alive)} tensor view(s) of the segment are still "
                                   "alive; delete them first (or close(force=True))")
            self._L.cpb_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        # a tensor view keeps the segment alive (its array interface owns a reference), so by the time
        # the segment is collected no view is left
        try:
            self.close(force=True)
        except Exception:
            pass
