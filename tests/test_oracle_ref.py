"""Pin the oracle against the REFERENCE's own code where that exists outside the Fortran tool chain:
src/cuuser_utils_kernels.cu (the helper kernels of the reference's cuFFT path: set_psi, density sum,
pointwise V*psi, phasen, putz/getz, pack/unpack) is compiled for the host from where it lies under
/root/reference (oracle/Makefile -> oracle/_ref/libcuuser_ref.so) and executed here.

What this pins: the packing rule and the nzhs/indzs maps as consumed by the reference's scatter, the
density and V*psi formulas, phasen's sign pattern, the z-band insertion, and the ray -> plane map
msp as consumed by the reference's unpack, the clearing of the padding - and, by chaining them with 1-D
DFTs in the order and with the transpositions of fftnew (fftmain_utils.mod.F90:92-104, 122-136), the whole
staged sparse transform against the oracle's dense one.  The 1-D DFT convention is the one the
reference's GPU path states in code: isign = +1 -> CUFFT_FORWARD, -1 -> CUFFT_INVERSE, then zdscal(scale)
(mltfft_utils.mod.F90:636-646).  What stays a restatement: the loop structure of vpsi / rhoofr."""
import numpy as np
import pytest

from oracle import cpmd_oracle as orc
from oracle import ref_cuuser as ref

pytestmark = pytest.mark.skipif(ref.load() is None, reason="oracle/_ref/libcuuser_ref.so not built "
                                "(needs /root/reference at build time)")


@pytest.fixture(scope="module", params=[16, (16, 20, 24), 30])
def geo(request):
    return orc.make_geometry(request.param)


def test_library_is_built_from_the_reference_tree():
    assert ref.load().ref_source().decode().endswith("src/cuuser_utils_kernels.cu")


def test_set_psi_matches_reference_kernels(geo):
    c0, _, _ = orc.synthetic_inputs(geo, 2)
    assert np.array_equal(orc.set_psi_2_states_g(geo, c0[0], c0[1]), ref.set_psi_2_states_g(geo, c0[0], c0[1]))
    assert np.array_equal(orc.set_psi_1_state_g(geo, c0[0]), ref.set_psi_1_state_g(geo, c0[0]))
    a = 0.3 - 1.1j                                           # complex alpha: reference kernel vs formula
    want = np.zeros(geo.kr[0] * geo.nrays, complex)
    want[geo.nzhs - 1] = a * c0[1]
    want[geo.indzs - 1] = a * np.conj(c0[1])   # reference: (ar cx + ai cy, -ar cy + ai cx) = alpha * conj(c)
    if geo.geq0:
        want[geo.nzhs[0] - 1] = a * c0[1, 0]
    assert np.abs(ref.set_psi_1_state_g(geo, c0[1], a) - want).max() < 1e-15


def test_density_sum_and_pointwise_match_reference_kernels(geo):
    rng = np.random.default_rng(1)
    psi = rng.standard_normal(geo.nnr1) + 1j * rng.standard_normal(geo.nnr1)
    rho0 = rng.random(geo.nnr1)
    got = ref.build_density_sum(0.7, 1.9, psi, rho0.copy())
    assert np.array_equal(got, rho0 + (0.7 * psi.real * psi.real + 1.9 * psi.imag * psi.imag)) or \
        np.abs(got - (rho0 + 0.7 * psi.real ** 2 + 1.9 * psi.imag ** 2)).max() < 1e-15
    v = rng.random(geo.nnr1)
    assert np.array_equal(ref.pointwise_cxr(psi, v), v * psi)


def test_phasen_matches_reference_kernel(geo):
    rng = np.random.default_rng(2)
    f = rng.standard_normal(geo.nnr1) + 1j * rng.standard_normal(geo.nnr1)
    got = ref.phasen(geo, f).reshape(geo.kr[2], geo.kr[1], geo.kr[0])
    want = orc.phasen(geo, f)
    n1, n2, n3 = geo.nr
    assert np.array_equal(got[:n3, :n2, :n1], want[:n3, :n2, :n1])
    # the reference touches only the true mesh; pads keep their value (they hold zeros in practice)
    assert np.array_equal(got[n3:], f.reshape(got.shape)[n3:])


def _mltfft_nt(a, ldax, n, m, inverse, scale=1.0):
    """mltfft('N','T',a,ldax,m,b,m,ldax,n,m,isign,scale) the way the reference's GPU path defines it
    (mltfft_cuda, mltfft_utils.mod.F90:612-655): a batched Z2Z DFT - isign = +1 -> CUFFT_FORWARD
    (e^{-i...}), otherwise CUFFT_INVERSE (e^{+i...}, unnormalised) (:636-641) - of the m columns a(1:n, j),
    written transposed to b(j, 1:n), then scaled (:644) and its padding cleared by the reference's own
    SetBlock2Zero kernel (:645).  The padding of b is filled with garbage first to see that kernel work."""
    a2 = a.reshape(m, ldax)[:, :n]                            # [transform j][element]
    t = (np.fft.ifft(a2, axis=1) * n) if inverse else np.fft.fft(a2, axis=1)
    b = np.full((ldax, m), 9.0 - 9.0j)                        # b(j, k) -> flat k*m + j
    b[:n, :] = (t * scale).T
    return ref.setblock2zero(b.reshape(-1), "T", n, m, m, ldax)


def _mltfft_tn(a, ldbx, n, m, scale=1.0):
    """mltfft('T','N',a,m,ldbx,b,ldbx,m,n,m,isign=+1,scale): input transposed a(m, ldbx), output b(ldbx, m)."""
    a2 = a.reshape(ldbx, m)[:n, :].T
    t = np.fft.fft(a2, axis=1) * scale
    b = np.full((m, ldbx), 9.0 - 9.0j)
    b[:, :n] = t
    return ref.setblock2zero(b.reshape(-1), "N", n, m, ldbx, m)


def test_staged_sparse_transforms_through_reference_kernels(geo):
    """fftnew(isign=-1, sparse) and fftnew(isign=+1, sparse) assembled from the reference's own data
    movement kernels + 1-D DFTs, against the oracle's dense transforms (and thereby against everything the
    GPU library is compared with)."""
    n1, n2, n3 = geo.nr
    kr1, kr2, kr3 = geo.kr
    nzb = geo.kr3max - geo.kr3min + 1
    c0, f, v = orc.synthetic_inputs(geo, 2)
    psi = ref.set_psi_2_states_g(geo, c0[0], c0[1])                                   # (kr1s, nrays)
    # ---- inverse (fftmain_utils.mod.F90:92-104)
    xf = _mltfft_nt(psi, kr1, n1, geo.nrays, True)                                    # xf(nrays, kr1s)
    yf = ref.unpack_x2y(geo, xf, kr1)                                                 # [x][zr][y]  (pack_x2y = copy)
    xf = _mltfft_nt(yf, kr2, n2, nzb * kr1, True)                                     # [y][x][zr]
    yf = ref.putz(xf, geo.kr3min, geo.kr3max, kr3, kr1 * kr2)                         # [y][x][z]
    out = _mltfft_nt(yf, kr3, n3, kr1 * kr2, True)                                    # [z][y][x]
    want = orc.invfftn_sparse(geo, orc.set_psi_2_states_g(geo, c0[0], c0[1]))
    assert np.abs(out - want).max() < 1e-12 * np.abs(want).max()
    # ---- V * psi, density (the two consumers)
    vp = ref.pointwise_cxr(out, v)
    rho = ref.build_density_sum(f[0], f[1], out, np.zeros(geo.nnr1))
    assert np.abs(rho - orc.rhoofr(geo, c0, f, 1.0, 1.0)["rhoe"]).max() < 1e-12 * np.abs(rho).max()
    # ---- forward (fftmain_utils.mod.F90:122-136): z, getz, y, pack_y2x, x with the 1/N scale
    m = kr1 * kr2
    xf = _mltfft_tn(vp, kr3, n3, m)                                                   # (qr3s, m): [jj][z]
    ff = ref.getz(xf, geo.kr3min, geo.kr3max, kr3, m)                                 # [jj][zr]
    # f(m' = nzb*qr1, qr2s) transposed input of the y pass: element (j = zr + nzb*x, y) at j + m'*y
    yf = _mltfft_tn(ff, kr2, n2, nzb * kr1)                                           # (qr2s, m'): [x][zr][y]
    xf = ref.pack_y2x(geo, yf, kr1)                                                   # [x][ray]   (unpack_y2x = copy)
    g = _mltfft_tn(xf, kr1, n1, geo.nrays, 1.0 / (n1 * n2 * n3))                      # (qr1s, nrays)
    want_g = orc.fwfftn_sparse(geo, v * want)
    assert np.abs(g - want_g).max() < 1e-12 * np.abs(want_g).max()


def test_dense_transforms_through_reference_kernels():
    """The dense branch of fftnew (fftmain_utils.mod.F90:105-120, 137-153) on the density-cutoff sphere,
    assembled from the reference's unpack/pack + phasen kernels and 1-D DFTs, against the oracle's
    invfftn_dense / fwfftn_dense (what cpb_vofrho_local is compared with).  msqf = y + (z-1)*kr2s
    (fftprp_utils.mod.F90:259-268, full set) is derived here from the ray table."""
    geo = orc.make_density_geometry((16, 20, 24))
    n1, n2, n3 = geo.nr
    kr1, kr2, kr3 = geo.kr
    yy, zz = np.nonzero(geo.mg)                                   # 0-based (y, z) of every ray
    msqf = np.zeros(geo.nrays, dtype=np.int32)
    msqf[geo.mg[yy, zz] - 1] = (yy + 1) + zz * kr2
    sp8 = np.array([geo.nrays], dtype=np.int32)
    L = ref.load()
    rng = np.random.default_rng(4)
    vg = (rng.standard_normal(geo.ngw) + 1j * rng.standard_normal(geo.ngw)) * np.exp(-geo.hg / 30)
    vg[0] = vg[0].real
    v = np.zeros(kr1 * geo.nrays, complex)                        # vofrhob_utils.mod.F90:155-173
    v[geo.indzs - 1] = np.conj(vg)
    v[geo.nzhs - 1] = vg
    # ---- inverse: x, unpack (full planes), y, z, phasen
    xf = _mltfft_nt(v, kr1, n1, geo.nrays, True)
    mm = kr2 * kr3
    yf = np.zeros(mm * kr1, complex)
    L.ref_unpack_x2y(xf.ctypes.data, yf.ctypes.data, mm, kr1, geo.nrays * kr1, msqf.ctypes.data, geo.nrays,
                     sp8.ctypes.data, 0, 1)
    xf = _mltfft_nt(yf, kr2, n2, kr3 * kr1, True)                 # [y][x][z]
    out = _mltfft_nt(xf, kr3, n3, kr1 * kr2, True)                # [z][y][x]
    out = ref.phasen(geo, out)
    want = orc.invfftn_dense(geo, v)
    assert np.abs(out - want).max() < 1e-12 * np.abs(want).max()
    assert np.abs(out.imag).max() < 1e-12 * np.abs(want).max()    # a real potential
    # ---- forward: phasen, z, y, pack, x with the 1/N scale
    rho = out.real.copy()
    f = ref.phasen(geo, rho.astype(complex))
    xf = _mltfft_tn(f, kr3, n3, kr1 * kr2)                        # [jj][z]
    yf = _mltfft_tn(xf, kr2, n2, kr3 * kr1)                       # [x][z][y]
    xr = np.zeros(geo.nrays * kr1, complex)
    L.ref_pack_y2x(xr.ctypes.data, yf.ctypes.data, mm, kr1, geo.nrays * kr1, msqf.ctypes.data, geo.nrays,
                   sp8.ctypes.data, 0, 1)
    g = _mltfft_tn(xr, kr1, n1, geo.nrays, 1.0 / (n1 * n2 * n3))
    assert np.abs(g - orc.fwfftn_dense(geo, rho)).max() < 1e-13
    assert np.abs(g[geo.nzhs - 1] - vg).max() < 1e-12             # round trip on the sphere
