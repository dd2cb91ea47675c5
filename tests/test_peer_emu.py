"""cpb_peer_* (collectives over peer memory) on the kernel simulator: the ranks are threads of one
process whose "device" segments are host memory, so the slice arithmetic, the two-shot all-reduce,
the two-phase broadcast and the flag barriers are exercised with real concurrency on the CPU.  The
NVLink/IPC path itself is covered by tests/test_gpu_multi.py on a multi-GPU box."""
import threading

import numpy as np
import pytest

from cpmd_b200.dist import PeerSegment


class _Exchange:
    """all-gather of the handles between threads"""

    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.bar = threading.Barrier(world)

    def make(self, rank):
        def ex(b):
            self.slots[rank] = b
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()                       # nobody overwrites a slot before everyone copied
            return out
        return ex


def _run(world, fn):
    ex = _Exchange(world)
    err = []

    def body(r):
        try:
            fn(r, ex.make(r))
        except Exception as e:  # noqa: BLE001
            err.append(e)
            try:
                ex.bar.abort()
            except Exception:
                pass

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not err, err


@pytest.mark.parametrize("world,n", [(2, 1000), (3, 4098), (4, 17 * 17 * 17 + 1), (5, 14), (8, 50), (9, 130), (13, 40),
                                     (16, 300)])
def test_allreduce_and_bcast_threads_as_ranks(emu_cdll, world, n):
    n += n & 1
    rng = np.random.default_rng(world)
    data = rng.standard_normal((world, n))
    want = data[0].copy()
    for q in range(1, world):
        want = want + data[q]                     # the kernel's fixed rank order
    vsrc = rng.standard_normal(n)
    results = [None] * world
    done = threading.Barrier(world)

    def rank_fn(r, exchange):
        seg = PeerSegment(2 * n + 6, r, world, exchange=exchange, _cdll=emu_cdll)
        a = seg.numpy(0, n)
        v = seg.numpy(n + 4, n)
        a[:] = data[r]
        v[:] = vsrc if r == 1 % world else -7.0
        for _ in range(3):                        # repeated collectives: the epoch keeps increasing
            a[:] = data[r]
            seg.allreduce(0, n)
            assert np.array_equal(a, want)
        seg.bcast(n + 4, n, src=1 % world)
        seg.check()
        assert np.array_equal(v, vsrc)
        guard = seg.numpy(n, 4)                   # words between the two arrays stay untouched
        results[r] = a.copy()
        seg.barrier()
        done.wait()                               # nobody frees memory a peer may still read
        del a, v, guard
        seg.close()

    _run(world, rank_fn)
    for r in range(world):
        assert np.array_equal(results[r], want)


def test_argument_checks(emu_cdll):
    seg = PeerSegment(64, 0, 1, _cdll=emu_cdll)
    with pytest.raises(RuntimeError):
        seg.allreduce(1, 10)                      # odd offset
    with pytest.raises(RuntimeError):
        seg.allreduce(0, 128)                     # outside the segment
    with pytest.raises(RuntimeError):
        seg.bcast(0, 64, src=3)
    a = seg.numpy(0, 64)
    a[:] = 2.5
    seg.allreduce(0, 64)                          # world 1: identity
    seg.bcast(0, 64, src=0)
    assert np.all(a == 2.5)
    seg.close()


@pytest.mark.parametrize("world,nstate,ld", [(2, 5, 7), (3, 10, 33), (4, 3, 16), (8, 64, 5)])
def test_redist_c2_and_scalars(emu_cdll, world, nstate, ld):
    """cp_grp_redist(C2_vpsi) as an all-gather of the part_1d state blocks (vpsi_utils.mod.F90:708-712;
    blocks of 0 states when nstate < world) and the rank-ordered sum of the group-partial scalars."""
    from cpmd_b200.dist import state_block
    rng = np.random.default_rng(world + nstate)
    full = rng.standard_normal((nstate, ld)) + 1j * rng.standard_normal((nstate, ld))
    scal = rng.standard_normal((world, 3))
    off = 6
    done = threading.Barrier(world)

    def rank_fn(r, exchange):
        seg = PeerSegment(off + 2 * nstate * ld + 2, r, world, exchange=exchange, _cdll=emu_cdll)
        c2 = seg.numpy(off, 2 * nstate * ld).view(np.complex128).reshape(nstate, ld)
        c2[:] = -99.0                                  # foreign blocks hold garbage before the exchange
        first, cnt = state_block(nstate, r, world)
        c2[first:first + cnt] = full[first:first + cnt]
        for rep in range(3):
            seg.redist_c2(off, ld, nstate)
            got = seg.allreduce_scalars(scal[r] * (rep + 1))
            want = scal[0] * (rep + 1)
            for q in range(1, world):
                want = want + scal[q] * (rep + 1)      # the kernel's fixed rank order: bit-identical
            assert np.array_equal(np.array(got), want)
        seg.check()
        assert np.array_equal(c2, full)
        assert np.all(seg.numpy(0, off) == 0.0)        # words in front of the array untouched
        seg.barrier()
        done.wait()
        del c2
        seg.close()

    _run(world, rank_fn)


def test_missing_rank_raises_everywhere_and_leaves_data_alone(emu_cdll):
    """a rank that does not show up must not produce half-summed data.  Rank 2 of 3 never
    calls the collective: the others time out, the all-reduce kernels skip their work, check() raises
    on every rank - also on the late one, whose error word the waiting ranks raised."""
    world, n = 3, 64
    late = threading.Event()
    outcome = {}

    def rank_fn(r, exchange):
        seg = PeerSegment(n, r, world, exchange=exchange, _cdll=emu_cdll)
        seg.set_timeout_ms(50.0)
        a = seg.numpy(0, n)
        a[:] = float(r + 1)
        if r == 2:
            late.wait(timeout=60)                      # arrives after the others gave up
        seg.allreduce(0, n)
        try:
            seg.check()
            outcome[r] = "ok"
        except RuntimeError as e:
            outcome[r] = str(e)
        if r != 2:
            assert np.all(a == float(r + 1))           # untouched, not half-summed
            if all(q in outcome for q in (0, 1)):
                late.set()
        del a

    _run(world, rank_fn)
    late.set()
    assert all("timeout" in outcome[r] for r in range(world)), outcome


def test_close_refuses_while_views_are_alive(emu_cdll):
    seg = PeerSegment(64, 0, 1, _cdll=emu_cdll)

    class _T:                                          # stand-in for a tensor handed out by tensor()
        pass
    import weakref
    t = _T()
    seg._views.append(weakref.ref(t))
    with pytest.raises(RuntimeError):
        seg.close()
    del t
    seg.close()
