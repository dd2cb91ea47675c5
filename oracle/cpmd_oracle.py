"""CPU oracle for CPMD's Gamma-point ``vpsi`` + ``rhoofr`` hot path (NumPy/SciPy, FP64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be imported
only by ``tests/``, by ``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``, and always as the *checker*, never as the thing that is shipped
or measured as the GPU path.

PARITY PINNED BY REFERENCE CODE (three ways; what is left unpinned is listed at the end).  The reference tree
(/root/reference, CPMD 4.3) ships no golden vectors, known-answer tests or fixtures for this path, and as a whole it
cannot be compiled in the authoring container (Fortran 2008 + FFTW + MPI; no Fortran compiler is installed).
(1) The helper kernels of the reference's cuFFT code path, src/cuuser_utils_kernels.cu (set_psi_1/2_states_g,
build_density_sum, the pointwise V*psi, phasen, putz/getz via MatMov/Zeroing, pack/unpack x2y/y2x), compile from
their own file: oracle/Makefile builds them for the host from where they lie (nothing copied; CPU stand-ins for
the few CUDA names in oracle/ref_shim/) into oracle/_ref/libcuuser_ref.so, and tests/test_oracle_ref.py (a)
compares this file's functions with those kernels one by one and (b) assembles fftnew's staged sparse inverse and
forward transforms from the reference's data-movement kernels plus 1-D DFTs and checks them against the dense
transforms used below.  That pins the packing rule, the nzhs/indzs and msp index maps as the reference consumes
them, the z-band insertion, phasen and the density / V*psi formulas.  (2) The same sources plus cuuser_utils.cu
compile with nvcc: oracle/ref_gpu_driver.cu runs them with cuFFT plans laid out like cp_cufft_utils and the stage
order of fftcu_methods on the GPU (tests/test_gpu_reference_arm.py compares the library with it).  (3) The parts
that are plain Fortran loops - the pairing loops, the occupation rules, the +-G unpack with the kinetic term and the
-f/2 scale, the density coefficients and accumulation, kin_energy, dotp, the LSD branches and post-processing,
rhoofr_c's k-point loop and the k-point unpack, tauofr / vtaupsi (dpsisc, tauadd, taupot, ftauadd), ppener, hfxab /
hfxaa - are EXECUTED
from the reference's own statements: oracle/fsnip.py reads the cited line ranges of vpsi_utils / rhoofr_utils /
density_utils / kin_energy_utils / dotp_utils / part_1d from /root/reference/src and runs them statement by
statement on NumPy data; tests/test_fsnip_pin.py compares this file with them (live, where the tree exists) and
with the fixtures they produced (tests/golden/fsnip, tools/make_golden_fsnip.py), which the staged C oracle, the
simulator build of the kernels and the GPU tests are checked against too.
Still a restatement (pinned by known-answer tests derived from the reference's formulas, tests/test_oracle.py, and
by the independent second restatement oracle/staged_oracle.c): the 1-D DFT itself (its sign/scale convention is
the one mltfft_cuda states in code: isign = +1 -> CUFFT_FORWARD, -1 -> CUFFT_INVERSE, then zdscal(scale),
mltfft_utils.mod.F90:636-646 - and the cuFFT run of (2) confirms it) and hfx_old's outer pair loop.

All "Fortran" indices kept in arrays here are 1-based exactly like the reference's (``inyh``,
``nzhs``, ``indzs``); they are converted at the point of use.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.fft as sfft

# ----------------------------------------------------------------------------------------------
# mesh helpers
# ----------------------------------------------------------------------------------------------

#: admissible FFT lengths, roots 2,3,5,7 (fftchk_utils.mod.F90:75-97, first part of the LFT table)
LFT = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20, 21, 24, 25, 27, 28, 30, 32, 35, 36, 40,
       42, 45, 48, 49, 50, 54, 56, 60, 63, 64, 70, 72, 75, 80, 81, 84, 90, 96, 98, 100, 105, 108,
       112, 120, 125, 126, 128, 135, 140, 144, 147, 150, 160, 162, 168, 175, 180, 189, 192, 196,
       200, 210, 216, 224, 225, 240, 243, 245, 250, 252, 256, 270, 280, 288, 294, 300, 315, 320,
       324, 336, 343, 350, 360, 375, 378, 384, 392, 400, 405, 420, 432, 441, 448, 450, 480, 486,
       490, 500, 504, 512]


def fftchk(m: int, n: int = 2) -> int:
    """Next admissible FFT length >= m (n=1) / next even one (n=2); fftchk_utils.mod.F90:60-131."""
    for v in LFT:
        if v >= m and (n == 1 or v % 2 == 0):
            return v
    raise ValueError("mesh too large for table")


def leadim(nr: int) -> int:
    """Odd-padded leading dimension kr = nr + MOD(nr+1,2); loadpa_utils.mod.F90:509-525."""
    return nr + (nr + 1) % 2


# ----------------------------------------------------------------------------------------------
# G vectors (loadpa) and FFT index maps (fftprp)
# ----------------------------------------------------------------------------------------------

@dataclass
class Geometry:
    nr: tuple          # (nr1s, nr2s, nr3s)
    kr: tuple          # (kr1s, kr2s, kr3s) padded leading dimensions
    ngw: int
    inyh: np.ndarray   # (3, ngw) int32, 1-based box position of +G  (loadpa_utils.mod.F90:306-308)
    hg: np.ndarray     # (ngw,) |G|^2 in units of tpiba2              (rggen_utils.mod.F90:121-129)
    geq0: bool         # first vector is G=0                          (loadpa_utils.mod.F90:414-420)
    nrays: int         # ngrays = msrays
    mg: np.ndarray     # (kr2s, kr3s) int32 ray number (1-based) or 0 (fftprp_utils.mod.F90:209-217)
    nzhs: np.ndarray   # (ngw,) int32 1-based index of +G into ray storage (kr1s, nrays)
    indzs: np.ndarray  # (ngw,) int32 1-based index of -G
    kr3min: int        # 1-based z band (fftprp_utils.mod.F90:177-192)
    kr3max: int
    msp2: np.ndarray   # (nrays,) ray -> y + (z-kr3min)*kr2s, 1-based   (fftprp_utils.mod.F90:259-268)

    @property
    def nnr1(self):
        return self.kr[0] * self.kr[1] * self.kr[2]


def gvectors(nr, gcutw, b=None):
    """Half-sphere G enumeration of ``loadpa`` (loadpa_utils.mod.F90:282-335) restricted to the
    wavefunction cutoff |G|^2 < gcutw, followed by the |G|^2 sort (:408, gsort :748-792).

    Enumeration rule: i>=0; i==0 => j>=0; i==j==0 => k>=0.  ``inyh = nh + (i,j,k)``,
    ``nh = nr/2+1``.  The order inside a |G|^2 shell is an implementation detail of the
    reference's sort (symmetry broken by a sqrt(ig-1)*eps perturbation); neither vpsi nor rhoofr
    depend on it, so a stable sort on the exact |G|^2 (enumeration order inside a shell) is used,
    which puts G=0 first exactly like the reference.
    Returns (inyh(3,ngw) int32 1-based, hg(ngw) float64).
    """
    nr1, nr2, nr3 = nr
    if b is None:
        b = np.eye(3)
    b = np.asarray(b, dtype=np.float64)
    nh = (nr1 // 2 + 1, nr2 // 2 + 1, nr3 // 2 + 1)
    i = np.arange(0, nr1)
    j = np.arange(-nr2 + 1, nr2)
    k = np.arange(-nr3 + 1, nr3)
    # limit the candidate ranges by a bounding box (pure speed-up, same set)
    bn = np.linalg.norm(np.linalg.inv(b), axis=0)  # |G_i| >= |i| / |a_i|-ish bound; conservative
    rad = np.sqrt(gcutw)
    i = i[i <= rad * bn[0] + 1]
    j = j[np.abs(j) <= rad * bn[1] + 1]
    k = k[np.abs(k) <= rad * bn[2] + 1]
    I, J, K = np.meshgrid(i, j, k, indexing="ij")  # enumeration order: i outer, j, k inner
    I = I.ravel(); J = J.ravel(); K = K.ravel()
    keep = (I > 0) | ((I == 0) & (J > 0)) | ((I == 0) & (J == 0) & (K >= 0))
    I, J, K = I[keep], J[keep], K[keep]
    t = I[:, None] * b[0][None, :] + J[:, None] * b[1][None, :] + K[:, None] * b[2][None, :]
    g2 = (t * t).sum(axis=1)
    sel = g2 < gcutw
    I, J, K, g2 = I[sel], J[sel], K[sel], g2[sel]
    order = np.argsort(g2, kind="stable")
    I, J, K, g2 = I[order], J[order], K[order], g2[order]
    inyh = np.stack([nh[0] + I, nh[1] + J, nh[2] + K]).astype(np.int32)
    # box must contain both +G and -G without touching index 1 whose mirror falls outside
    for d in range(3):
        if inyh[d].min() < 2 or inyh[d].max() > nr[d]:
            raise ValueError("cutoff sphere does not fit the mesh (dual < 4?)")
    return inyh, g2.astype(np.float64)


def fft_maps(nr, inyh, hg) -> Geometry:
    """``fftprp_default_init`` for one rank per group (fftprp_utils.mod.F90:139-285)."""
    nr1, nr2, nr3 = nr
    kr = (leadim(nr1), leadim(nr2), leadim(nr3))
    nh1, nh2, nh3 = nr1 // 2 + 1, nr2 // 2 + 1, nr3 // 2 + 1
    ngw = inyh.shape[1]
    ny1, ny2, ny3 = (inyh[0].astype(np.int64), inyh[1].astype(np.int64), inyh[2].astype(np.int64))
    iny1, iny2, iny3 = 2 * nh1 - ny1, 2 * nh2 - ny2, 2 * nh3 - ny3          # :147-148, :272-277
    mg = np.zeros((kr[1] + 1, kr[2] + 1), dtype=np.int64)                   # 1-based, [0] unused
    mg[ny2, ny3] = 1
    mg[iny2, iny3] = 1                                                      # :149-150
    mz = np.zeros(kr[2] + 2, dtype=np.int64)
    mz[ny3] = 1
    mz[iny3] = 1
    zs = np.nonzero(mz)[0]
    kr3min, kr3max = int(zs.min()), int(zs.max())                           # :177-192
    # ray numbering: z outer, y inner (:209-217)
    img = 0
    for jz in range(1, kr[2] + 1):
        ys = np.nonzero(mg[1:, jz])[0] + 1
        for iy in ys:
            img += 1
            mg[iy, jz] = img
    nrays = img
    nzhs = ny1 + (mg[ny2, ny3] - 1) * kr[0]                                 # :278
    indzs = iny1 + (mg[iny2, iny3] - 1) * kr[0]                             # :279
    msp2 = np.zeros(nrays, dtype=np.int64)
    yy, zz = np.nonzero(mg)
    msp2[mg[yy, zz] - 1] = yy + (zz - kr3min) * kr[1]                       # :259-268
    geq0 = bool(hg[0] < 1.0e-5)                                             # loadpa :414-420
    return Geometry(nr=tuple(nr), kr=kr, ngw=ngw, inyh=inyh, hg=hg, geq0=geq0, nrays=nrays,
                    mg=mg[1:, 1:].astype(np.int32), nzhs=nzhs.astype(np.int32),
                    indzs=indzs.astype(np.int32), kr3min=kr3min, kr3max=kr3max,
                    msp2=msp2.astype(np.int32))


def make_geometry(nr, gcutw=None, b=None) -> Geometry:
    """Convenience: synthetic cubic-cell geometry with dual 4: gcutw = (n/4)^2."""
    if isinstance(nr, int):
        nr = (nr, nr, nr)
    if gcutw is None:
        gcutw = (min(nr) / 4.0) ** 2
    inyh, hg = gvectors(nr, gcutw, b)
    return fft_maps(nr, inyh, hg)


# ----------------------------------------------------------------------------------------------
# state-group decomposition (part_1d)
# ----------------------------------------------------------------------------------------------

def part_1d_nbr_el_in_blk(n_elem, proc, nproc):
    """part_1d.mod.F90:22-38."""
    res = n_elem % nproc
    nbr = (n_elem - res) // nproc
    return nbr + 1 if proc < res else nbr


def part_1d_get_el_in_blk(i_elem, n_elem, proc, nproc):
    """1-based element -> 1-based global index; part_1d.mod.F90:41-57."""
    res = n_elem % nproc
    nbr = (n_elem - res) // nproc
    return i_elem + nbr * proc + min(proc, res)


def state_pairs(nstate, group=0, ngroups=1):
    """Pairs (is1, is2) (0-based; is2 = None for a trailing single state) formed inside the
    group's block exactly like the hot loops (vpsi_utils.mod.F90:377-383,
    rhoofr_utils.mod.F90:306-310)."""
    nblk = part_1d_nbr_el_in_blk(nstate, group, ngroups)
    out = []
    for i in range(1, nblk + 1, 2):
        is1 = part_1d_get_el_in_blk(i, nstate, group, ngroups)
        is2 = part_1d_get_el_in_blk(i + 1, nstate, group, ngroups) if i + 1 <= nblk else None
        out.append((is1 - 1, None if is2 is None else is2 - 1))
    return out


# ----------------------------------------------------------------------------------------------
# packing, transforms
# ----------------------------------------------------------------------------------------------

def set_psi_2_states_g(geo: Geometry, c1, c2):
    """psi(nzfs)=c1+i*c2, psi(inzs)=conj(c1)+i*conj(c2), G=0 rewritten last
    (state_utils.mod.F90:171-189).  Returns ray storage psi(kr1s*nrays)."""
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    psi[geo.nzhs - 1] = c1 + 1j * c2
    psi[geo.indzs - 1] = np.conj(c1) + 1j * np.conj(c2)
    if geo.geq0:
        psi[geo.nzhs[0] - 1] = c1[0] + 1j * c2[0]
    return psi


def set_psi_1_state_g(geo: Geometry, c1, alpha=1.0):
    """state_utils.mod.F90:132-168."""
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    psi[geo.nzhs - 1] = alpha * c1
    psi[geo.indzs - 1] = alpha * np.conj(c1)
    if geo.geq0:
        psi[geo.nzhs[0] - 1] = alpha * c1[0]
    return psi


def _rays_to_box(geo: Geometry, psi_rays):
    """Ray storage (kr1s, nrays) -> dense box [z, y, x] of the true mesh size (zeros elsewhere).
    Ray r sits at (y,z) with mg(y,z)=r (fftprp_utils.mod.F90:209-217)."""
    n1, n2, n3 = geo.nr
    rays = psi_rays.reshape(geo.nrays, geo.kr[0])
    box = np.zeros((n3, n2, n1), dtype=np.complex128)
    yy, zz = np.nonzero(geo.mg[:n2, :n3])
    r = geo.mg[yy, zz] - 1
    box[zz, yy, :] = rays[r, :n1]
    return box


def _box_to_rays(geo: Geometry, box):
    n1, n2, n3 = geo.nr
    rays = np.zeros((geo.nrays, geo.kr[0]), dtype=np.complex128)
    yy, zz = np.nonzero(geo.mg[:n2, :n3])
    r = geo.mg[yy, zz] - 1
    rays[r, :n1] = box[zz, yy, :]
    return rays.reshape(-1)


def invfftn_sparse(geo: Geometry, psi_rays):
    """``invfftn(psi,.TRUE.)`` = fftnew(isign=-1, sparse) (fftmain_utils.mod.F90:373-401, 92-104):
    unnormalised e^{+i...} transform, box index as frequency (no phase factor in the sparse
    branch), result in the padded real-space layout (kr1, kr2s, kr3s), x fastest, pads zero
    (mltfft_utils.mod.F90:227-253).  The dense 3-D transform of the zero-filled box is the same
    linear map as the staged x/y/z passes."""
    n1, n2, n3 = geo.nr
    box = _rays_to_box(geo, psi_rays)
    r = sfft.ifftn(box, norm="forward", workers=-1)      # "forward" => no 1/N on the inverse
    out = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]), dtype=np.complex128)
    out[:n3, :n2, :n1] = r
    return out.reshape(-1)


def fwfftn_sparse(geo: Geometry, psi_r):
    """``fwfftn(psi,.TRUE.)`` = fftnew(isign=+1, sparse): e^{-i...} with scale 1/(n1 n2 n3)
    applied in the last pass (fftmain_utils.mod.F90:403-431, 122-136); returns ray storage."""
    n1, n2, n3 = geo.nr
    box = psi_r.reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n3, :n2, :n1]
    g = sfft.fftn(box, norm="forward", workers=-1)       # "forward" => 1/N on the forward
    return _box_to_rays(geo, g)


# ----------------------------------------------------------------------------------------------
# G-space reductions
# ----------------------------------------------------------------------------------------------

def dotp(geo: Geometry, a, b):
    """dotp_utils.mod.F90:26-53 — half-sphere weight 2, G=0 weight 1."""
    if geo.geq0:
        d = a[0].real * b[0].real
    else:
        d = 2.0 * (a[0].real * b[0].real + a[0].imag * b[0].imag)
    if a.shape[0] > 1:
        d += 2.0 * float(np.dot(a[1:].real, b[1:].real) + np.dot(a[1:].imag, b[1:].imag))
    return float(d)


def kin_energy(geo: Geometry, c0, f, tpiba2):
    """kin_energy_utils.mod.F90:62-110 (akin == 0 branch).  c0 is (nstate, ngw) here (each row one
    Fortran column c0(:,i)).  Returns (ekin, rsum)."""
    rsum = 0.0
    xkin = 0.0
    for i in range(c0.shape[0]):
        if f[i] != 0.0:
            rsum += f[i] * dotp(geo, c0[i], c0[i])
            sk1 = float(np.sum(geo.hg * (c0[i].real ** 2 + c0[i].imag ** 2)))
            xkin += f[i] * sk1
    return xkin * tpiba2, rsum


# ----------------------------------------------------------------------------------------------
# the two hot routines
# ----------------------------------------------------------------------------------------------

def rhoofr(geo: Geometry, c0, f, omega, tpiba2, group=0, ngroups=1):
    """``rhoofr`` (rhoofr_utils.mod.F90:122-644), Gamma point, no LSD/LSE/tau/double grid.

    c0: (nstate, ngw) complex128.  Returns dict(rhoe (nnr1,), ekin, rsum_g, rsum_r) where rhoe is
    the *group-partial* density when ngroups>1 (the caller sums over groups, cp_grp_redist
    :457-461) and ekin/rsum_g are computed over all states as the reference does (:178)."""
    nstate = c0.shape[0]
    ekin, rsum = kin_energy(geo, c0, f, tpiba2)                       # :178
    rhoe = np.zeros(geo.nnr1, dtype=np.float64)                       # :198
    for is1, is2 in state_pairs(nstate, group, ngroups):              # :306-310
        tfcal = f[is1] != 0.0 or (is2 is not None and f[is2] != 0.0)  # :312-316
        if not tfcal:
            continue
        if is2 is None:
            psi = set_psi_1_state_g(geo, c0[is1])                     # :329-330
        else:
            psi = set_psi_2_states_g(geo, c0[is1], c0[is2])           # :332
        psi = invfftn_sparse(geo, psi)                                # :346
        coef3 = f[is1] / omega                                        # :369
        coef4 = 0.0 if is2 is None else f[is2] / omega                # :370-374
        rhoe += coef3 * psi.real ** 2 + coef4 * psi.imag ** 2         # density_utils :61-83
    n1, n2, n3 = geo.nr
    rsum1 = float(np.sum(rhoe)) * omega / float(n1 * n2 * n3)         # :607-619
    return dict(rhoe=rhoe, ekin=ekin, rsum_g=rsum, rsum_r=rsum1)


def vpsi(geo: Geometry, c0, c2, f, vpot, tpiba2, group=0, ngroups=1, redist_c2=False,
         tksham=False):
    """``vpsi`` (vpsi_utils.mod.F90:120-732), Gamma point, RKS, akin == 0.

    c0, c2: (nstate, ngw) complex128 (c2 is accumulated into: ``c2 += C2_vpsi`` :717);
    vpot: (nnr1,) padded real-space potential.  Only the group's block of states is touched.
    Returns the new c2 (a copy)."""
    nstate = c0.shape[0]
    c2v = np.zeros_like(c2)                                           # :199-203
    for is1, is2 in state_pairs(nstate, group, ngroups):              # :376-383
        if is2 is None:
            psi = set_psi_1_state_g(geo, c0[is1])                     # :432-433
        else:
            psi = set_psi_2_states_g(geo, c0[is1], c0[is2])           # :435
        psi = invfftn_sparse(geo, psi)                                # :443
        psi = vpot * psi                                              # :487-493
        psi = fwfftn_sparse(geo, psi)                                 # :552
        fi = f[is1] * 0.5                                             # :627
        if fi == 0.0:
            fi = 0.5 if tksham else 1.0                               # :628-629
        fip1 = 0.0
        if is2 is not None:
            fip1 = f[is2] * 0.5                                       # :631
        if fip1 == 0.0:
            fip1 = 0.5 if tksham else 1.0                             # :632-633
        psin = psi[geo.nzhs - 1]                                      # :662
        psii = psi[geo.indzs - 1]                                     # :663
        fp = psin + psii
        fm = psin - psii
        g2 = tpiba2 * geo.hg
        c2v[is1] = -fi * (g2 * c0[is1] + (fp.real + 1j * fm.imag))    # :666-667
        if is2 is not None:
            c2v[is2] = -fip1 * (g2 * c0[is2] + (fp.imag - 1j * fm.real))   # :668-670
    return c2 + c2v                                                   # :717


# ----------------------------------------------------------------------------------------------
# LSD (cntl%tlsd) variants: states 1..nsup are alpha, nsup+1..nstate beta (spin_mod%nsup)
# ----------------------------------------------------------------------------------------------

def lsd_finish(geo: Geometry, rhoe2, omega):
    """rhoofr_utils.mod.F90:543-559 + :607-619, applied to the two channel densities *after* the
    group sum (cp_grp_redist, :457-461): csums / csumsabs from alpha - beta, then column 1 becomes
    alpha + beta; column 2 stays beta.  rhoe2: (2, nnr1), modified in place.
    Returns (rsum_r, csums, csumsabs)."""
    n1, n2, n3 = geo.nr
    w = omega / float(n1 * n2 * n3)
    d = rhoe2[0] - rhoe2[1]
    csums = float(np.sum(d)) * w
    csumsabs = float(np.sum(np.abs(d))) * w
    rhoe2[0] += rhoe2[1]
    return float(np.sum(rhoe2[0])) * w, csums, csumsabs


def rhoofr_lsd(geo: Geometry, c0, f, omega, tpiba2, nsup, group=0, ngroups=1):
    """``rhoofr`` with cntl%tlsd (rhoofr_utils.mod.F90:375-385): the Re part of a pair goes to the
    spin channel of is1, the Im part to that of is2 (a missing partner counts as is2 = nstate+1,
    :308, i.e. beta, with coef4 = 0).  Returns dict(rhoe (2, nnr1), ekin, rsum_g, rsum_r, csums,
    csumsabs): for ngroups == 1 rhoe[0] = alpha+beta, rhoe[1] = beta (:543-559); for ngroups > 1
    the group's *partial* alpha and beta densities (the reference sums over groups first; finish
    with :func:`lsd_finish`), and rsum_r/csums/csumsabs are None."""
    nstate = c0.shape[0]
    ekin, rsum = kin_energy(geo, c0, f, tpiba2)
    rhoe = np.zeros((2, geo.nnr1), dtype=np.float64)
    for is1, is2 in state_pairs(nstate, group, ngroups):
        tfcal = f[is1] != 0.0 or (is2 is not None and f[is2] != 0.0)
        if not tfcal:
            continue
        psi = set_psi_1_state_g(geo, c0[is1]) if is2 is None else set_psi_2_states_g(geo, c0[is1], c0[is2])
        psi = invfftn_sparse(geo, psi)
        coef3 = f[is1] / omega
        coef4 = 0.0 if is2 is None else f[is2] / omega
        ispin1 = 1 if (is1 + 1) > nsup else 0                         # :378 (1-based is1)
        ispin2 = 1 if ((nstate + 1) if is2 is None else (is2 + 1)) > nsup else 0   # :379
        if ispin1 == ispin2:
            rhoe[ispin1] += coef3 * psi.real ** 2 + coef4 * psi.imag ** 2      # :381
        else:
            rhoe[ispin1] += coef3 * psi.real ** 2                              # :383
            rhoe[ispin2] += coef4 * psi.imag ** 2                              # :384
    out = dict(rhoe=rhoe, ekin=ekin, rsum_g=rsum, rsum_r=None, csums=None, csumsabs=None)
    if ngroups == 1:
        out["rsum_r"], out["csums"], out["csumsabs"] = lsd_finish(geo, rhoe, omega)
    return out


def vpsi_lsd(geo: Geometry, c0, c2, f, vpot2, tpiba2, nsup, group=0, ngroups=1, tksham=False):
    """``vpsi`` with cntl%tlsd and ispin = 2 (vpsi_utils.mod.F90:450-482): vpot2 is (2, nnr1)
    [alpha, beta]; the pair that straddles the spin boundary (is1 == nsup, 1-based) gets
    V_alpha * Re(psi) + i V_beta * Im(psi), every other pair the potential of is1's spin."""
    nstate = c0.shape[0]
    c2v = np.zeros_like(c2)
    for is1, is2 in state_pairs(nstate, group, ngroups):
        psi = set_psi_1_state_g(geo, c0[is1]) if is2 is None else set_psi_2_states_g(geo, c0[is1], c0[is2])
        psi = invfftn_sparse(geo, psi)
        if is1 + 1 == nsup:                                                    # :451
            psi = vpot2[0] * psi.real + 1j * vpot2[1] * psi.imag               # :466-469
        else:
            lspin = 1 if (is1 + 1) > nsup else 0                               # :475-476
            psi = vpot2[lspin] * psi                                           # :481
        psi = fwfftn_sparse(geo, psi)
        fi = f[is1] * 0.5
        if fi == 0.0:
            fi = 0.5 if tksham else 1.0
        fip1 = 0.0
        if is2 is not None:
            fip1 = f[is2] * 0.5
        if fip1 == 0.0:
            fip1 = 0.5 if tksham else 1.0
        psin = psi[geo.nzhs - 1]
        psii = psi[geo.indzs - 1]
        fp = psin + psii
        fm = psin - psii
        g2 = tpiba2 * geo.hg
        c2v[is1] = -fi * (g2 * c0[is1] + (fp.real + 1j * fm.imag))
        if is2 is not None:
            c2v[is2] = -fip1 * (g2 * c0[is2] + (fp.imag - 1j * fm.real))
    return c2 + c2v


# ----------------------------------------------------------------------------------------------
# dense transforms on the density cutoff and the local part of vofrho
# ----------------------------------------------------------------------------------------------

def make_density_geometry(nr, gcut=None, b=None) -> Geometry:
    """Geometry of the DENSITY cutoff sphere |G|^2 < gcut = dual * gcutw (numpw_utils.mod.F90:180-184;
    dual 4 with gcutw = (n/4)^2 gives gcut = (n/2)^2): the nhg vectors of ``loadpa`` in the same
    sort order, whose first ngw entries are the wavefunction sphere.  The maps named nzhs/indzs in
    the returned object are the reference's nzh/indz for this set (fftprp_utils.mod.F90:269-285),
    ``ngw`` is nhg."""
    if isinstance(nr, int):
        nr = (nr, nr, nr)
    if gcut is None:
        gcut = (min(nr) / 2.0) ** 2
    inyh, hg = gvectors(nr, gcut, b)
    return fft_maps(nr, inyh, hg)


def phasen(geo: Geometry, f):
    """phasen (fftutil_utils.mod.F90:479-503): f(i,j,k) *= pf(MOD(k+j+i+1,2)+1), pf = (+1,-1), with
    1-based i,j,k, i.e. (-1)^(x+y+z) in 0-based mesh coordinates.  f: padded (kr3,kr2,kr1) box."""
    kr1, kr2, kr3 = geo.kr
    z, y, x = np.ogrid[:kr3, :kr2, :kr1]
    return f.reshape(kr3, kr2, kr1) * (1.0 - 2.0 * ((x + y + z) & 1))


def fwfftn_dense(geo: Geometry, f_r):
    """``fwfftn(v,.FALSE.)`` = fftnew(isign=+1, dense) (fftmain_utils.mod.F90:137-153): phasen, then
    the three e^{-i...} passes with the scale 1/(n1 n2 n3) in the last one; returns ray storage
    (read with nzh/indz)."""
    n1, n2, n3 = geo.nr
    box = phasen(geo, np.asarray(f_r, dtype=np.complex128))[:n3, :n2, :n1]
    g = sfft.fftn(box, norm="forward", workers=-1)
    return _box_to_rays(geo, g)


def invfftn_dense(geo: Geometry, v_rays):
    """``invfftn(v,.FALSE.)`` = fftnew(isign=-1, dense) (fftmain_utils.mod.F90:105-120): unnormalised
    e^{+i...} passes, then phasen.  Returns the padded complex real-space array (nnr1,)."""
    n1, n2, n3 = geo.nr
    box = _rays_to_box(geo, v_rays)
    r = sfft.ifftn(box, norm="forward", workers=-1)
    out = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]), dtype=np.complex128)
    out[:n3, :n2, :n1] = r
    return phasen(geo, out).reshape(-1)


def rho_to_g(geo: Geometry, rhoe):
    """vofrhoa_utils.mod.F90:88-95 + ppener_utils.mod.F90:91: v = CMPLX(rhoe,0); fwfftn(v,.FALSE.);
    rhog(ig) = v(nzh(ig))."""
    return fwfftn_dense(geo, rhoe)[geo.nzhs - 1]


def g_to_r(geo: Geometry, vg):
    """vofrhob_utils.mod.F90:155-173: v = 0; v(indz) = CONJG(vg); v(nzh) = vg; G=0 rewritten;
    invfftn(v,.FALSE.).  Returns the complex padded array (its real part is the potential)."""
    v = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    v[geo.indzs - 1] = np.conj(vg)
    v[geo.nzhs - 1] = vg
    if geo.geq0:
        v[geo.nzhs[0] - 1] = vg[0]
    return invfftn_dense(geo, v)


def ppener(geo: Geometry, rhog, scg, eivps, eirop):
    """ppener (ppener_utils.mod.F90:23-108).  Returns (eh, ei, ee, eps, vploc, vtemp), the four
    sums complex like the reference's."""
    nhg = geo.ngw
    vtemp = np.empty(nhg, dtype=np.complex128)
    ig1 = 0
    eh = ei = ee = eps = 0.0 + 0.0j
    vploc = 0.0
    if geo.geq0:                                                       # :58-70
        vp = eivps[0]
        vploc = vp.real
        eps = 0.5 * vp * np.conj(rhog[0])
        rp = eirop[0]
        rhet = rhog[0]
        rg = rhet + rp
        eh = 0.5 * scg[0] * rg.real * rg.real + 0.0j
        ei = 0.5 * scg[0] * rp * rp
        ee = 0.5 * scg[0] * rhet * rhet
        vtemp[0] = scg[0] * rg
        ig1 = 1
    vp = eivps[ig1:]
    rp = eirop[ig1:]
    rhet = rhog[ig1:]
    rg = rhet + rp                                                     # :92
    vcg = scg[ig1:] * rg                                               # :95
    vtemp[ig1:] = vcg + vp                                             # :96
    eh = eh + np.sum(vcg * np.conj(rg))                                # :98
    ei = ei + np.sum(scg[ig1:] * rp * np.conj(rp))                     # :100
    ee = ee + np.sum(scg[ig1:] * rhet * np.conj(rhet))                 # :102
    eps = eps + np.sum(np.conj(rhet) * vp)                             # :103
    return eh, ei, ee, eps, vploc, vtemp


def vofrho_local(geo: Geometry, rhoe, scg, eivps, eirop):
    """The local (G-space electrostatic) part of vofrho: vofrhoa_utils.mod.F90:88-102 (density to G,
    ppener) and vofrhob_utils.mod.F90:155-173 (potential back to real space), without eextern,
    forces, stress and exchange-correlation.  Returns dict(v (nnr1,) real, rhog, vtemp, eh, ei, ee,
    eps, vploc)."""
    rhog = rho_to_g(geo, rhoe)
    eh, ei, ee, eps, vploc, vtemp = ppener(geo, rhog, scg, eivps, eirop)
    v = g_to_r(geo, vtemp)
    return dict(v=np.ascontiguousarray(v.real), v_imag_max=float(np.abs(v.imag).max()), rhog=rhog, vtemp=vtemp,
                eh=eh, ei=ei, ee=ee, eps=eps, vploc=vploc)


def synthetic_vofrho_inputs(geo: Geometry, tpiba2=1.0, omega=1.0, seed=None):
    """Synthetic G-space inputs of ppener on the density sphere: scg = 4 pi / (tpiba2 hg), 0 at G = 0
    (the periodic Coulomb kernel), eivps / eirop = smooth random structure-factor-like complex arrays
    (real at G = 0).  ``omega`` only scales eirop like a charge density."""
    n = geo.nr[0]
    if seed is None:
        seed = 4321 + n
    rng = np.random.default_rng(seed)
    hg = geo.hg
    scg = np.zeros(geo.ngw)
    nz = hg > 1.0e-8
    scg[nz] = 4.0 * np.pi / (tpiba2 * hg[nz])
    gc = float(hg.max()) + 1.0
    damp = np.exp(-hg / (0.1 * gc))
    eivps = (rng.standard_normal(geo.ngw) + 1j * rng.standard_normal(geo.ngw)) * damp * (-0.5)
    eirop = (rng.standard_normal(geo.ngw) + 1j * rng.standard_normal(geo.ngw)) * damp * (-0.05 / omega)
    if geo.geq0:
        eivps[0] = eivps[0].real
        eirop[0] = eirop[0].real
    return scg, eivps, eirop


# ----------------------------------------------------------------------------------------------
# k-points (tkpts%tkpnt): complex states, one per transform, c0(ngwk = 2 ngw, nstate) per k-point:
# +G components first, then the -G components
# ----------------------------------------------------------------------------------------------

def set_psi_1_state_g_kpts(geo: Geometry, c1, alpha=1.0):
    """state_utils.mod.F90:192-224: psi(nzhs) = c1(1:ngw), psi(indzs) = c1(ngw+1:2ngw), G=0 last."""
    ngw = geo.ngw
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    psi[geo.nzhs - 1] = alpha * c1[:ngw]
    psi[geo.indzs - 1] = alpha * c1[ngw:2 * ngw]
    if geo.geq0:
        psi[geo.nzhs[0] - 1] = alpha * c1[0]
    return psi


def rhoofr_kpt(geo: Geometry, c0, f, wk, hgkp, hgkm, omega, tpiba2, group=0, ngroups=1, rhoe=None):
    """One k-point of ``rhoofr_c`` (rhoofr_c_utils.mod.F90:112-180): c0 (nstate, 2 ngw), f = crge%f(:,ikk),
    wk = wk(ikk), hgkp/hgkm = |k+G|^2, |k-G|^2 (ngw each).  ``rhoe`` (nnr1,) is accumulated into if
    given (the reference zeroes it once before the k-point loop, :107).  Returns dict(rhoe, ekin,
    rsum_g) with this k-point's contributions to ener_com%ekin (:138,182) and chrg%csumg (:119)."""
    nstate = c0.shape[0]
    ngw = geo.ngw
    rsum = 0.0
    xkin = 0.0
    for i in range(nstate):                                                    # :117-140
        if f[i] != 0.0:
            rsum += wk * f[i] * float(np.sum(c0[i].real ** 2 + c0[i].imag ** 2))
            sk1 = float(np.sum(hgkp * np.abs(c0[i, :ngw]) ** 2 + hgkm * np.abs(c0[i, ngw:2 * ngw]) ** 2))
            xkin += 0.5 * wk * f[i] * sk1
    if rhoe is None:
        rhoe = np.zeros(geo.nnr1, dtype=np.float64)
    nblk = part_1d_nbr_el_in_blk(nstate, group, ngroups)
    for i in range(1, nblk + 1):                                               # :143-178
        is1 = part_1d_get_el_in_blk(i, nstate, group, ngroups) - 1
        if f[is1] == 0.0:
            continue
        psi = invfftn_sparse(geo, set_psi_1_state_g_kpts(geo, c0[is1]))        # :150-154
        coef3 = wk * f[is1] / omega                                            # :165
        rhoe += coef3 * psi.real ** 2 + coef3 * psi.imag ** 2                  # :173
    return dict(rhoe=rhoe, ekin=xkin * tpiba2, rsum_g=rsum)


def vpsi_kpt(geo: Geometry, c0, c2, f, hgkp, hgkm, vpot, tpiba2, group=0, ngroups=1):
    """``vpsi`` with tkpts%tkpnt for one k-point (vpsi_utils.mod.F90:238,377,432,487-493,562-564,
    614-625): njump = 1, fi = f (2 if zero), C2(ig) = -fi (tpiba2/2 hgkp c0(ig) + psi(nzhs)),
    C2(ig+ngw) = -fi (tpiba2/2 hgkm c0(ig+ngw) + psi(indzs)), C2(1+ngw) = 0 if geq0; c2 += C2."""
    nstate = c0.shape[0]
    ngw = geo.ngw
    c2v = np.zeros_like(c2)
    nblk = part_1d_nbr_el_in_blk(nstate, group, ngroups)
    for i in range(1, nblk + 1):
        is1 = part_1d_get_el_in_blk(i, nstate, group, ngroups) - 1
        psi = invfftn_sparse(geo, set_psi_1_state_g_kpts(geo, c0[is1]))
        psi = fwfftn_sparse(geo, vpot * psi)
        fi = f[is1]
        if fi == 0.0:
            fi = 2.0                                                           # :563-564
        fp = psi[geo.nzhs - 1]
        fm = psi[geo.indzs - 1]
        c2v[is1, :ngw] = -fi * (0.5 * tpiba2 * hgkp * c0[is1, :ngw] + fp)      # :619-620
        c2v[is1, ngw:2 * ngw] = -fi * (0.5 * tpiba2 * hgkm * c0[is1, ngw:2 * ngw] + fm)   # :621-622
        if geo.geq0:
            c2v[is1, ngw] = 0.0                                                # :625
    return c2 + c2v


def synthetic_kpt_inputs(geo: Geometry, nstate, kvec=(0.25, 0.1, -0.3), seed=None):
    """Complex k-point states c0 (nstate, 2 ngw) (c0[ngw] = 0 where geq0: the -G slot of G=0 is
    unused), hgkp/hgkm for k = kvec (units of 2 pi/alat, cubic cell), occupations and a potential."""
    n = geo.nr[0]
    if seed is None:
        seed = 2468 + n + 7 * nstate
    rng = np.random.default_rng(seed)
    ngw = geo.ngw
    nh = np.array([v // 2 + 1 for v in geo.nr])
    g = (geo.inyh - nh[:, None]).astype(np.float64)
    k = np.asarray(kvec, dtype=np.float64)[:, None]
    hgkp = ((g + k) ** 2).sum(axis=0)
    hgkm = ((g - k) ** 2).sum(axis=0)
    gcutw = (min(geo.nr) / 4.0) ** 2
    c0 = np.empty((nstate, 2 * ngw), dtype=np.complex128)
    for i in range(nstate):
        c = rng.standard_normal(2 * ngw) + 1j * rng.standard_normal(2 * ngw)
        c *= np.exp(-np.concatenate([hgkp, hgkm]) / (0.25 * gcutw))
        if geo.geq0:
            c[ngw] = 0.0
        c /= np.sqrt(np.sum(np.abs(c) ** 2))
        c0[i] = c
    f = np.full(nstate, 2.0)
    f[1::4] = 0.0
    f[2::5] = 1.0
    n1, n2, n3 = geo.nr
    v = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]))
    v[:n3, :n2, :n1] = -rng.random((n3, n2, n1))
    return c0, f, hgkp, hgkm, v.reshape(-1)


# ----------------------------------------------------------------------------------------------
# meta-GGA (cntl%ttau): tauofr / vtaupsi
# ----------------------------------------------------------------------------------------------

def gk_cartesian(geo: Geometry, b=None):
    """cppt gk(3, ngw): Cartesian components of G in units of tpiba (rggen_utils.mod.F90:121-129,
    gk(:,ig) = i b1 + j b2 + k b3); cubic cell by default.  Returned as (ngw, 3) C-order = Fortran (3, ngw)."""
    nh = np.array([v // 2 + 1 for v in geo.nr])
    ijk = (geo.inyh - nh[:, None]).astype(np.float64).T
    if b is None:
        return np.ascontiguousarray(ijk)
    return np.ascontiguousarray(ijk @ np.asarray(b, dtype=np.float64))


def dpsisc(geo: Geometry, gk, c1, c2, k):
    """tauofr_utils.mod.F90:113-137: psi(nzhs) = gk(k,:) (c1 + i c2), psi(indzs) = -gk(k,:) (conj c1 + i conj c2),
    psi(G=0) = 0; c2 = None for the single-state form."""
    g = gk[:, k]
    psi = np.zeros(geo.kr[0] * geo.nrays, dtype=np.complex128)
    if c2 is None:
        psi[geo.nzhs - 1] = g * c1
        psi[geo.indzs - 1] = -g * np.conj(c1)
    else:
        psi[geo.nzhs - 1] = g * (c1 + 1j * c2)
        psi[geo.indzs - 1] = -g * (np.conj(c1) + 1j * np.conj(c2))
    if geo.geq0:
        psi[geo.nzhs[0] - 1] = 0.0
    return psi


def tauofr(geo: Geometry, c0, f, gk, omega, tpiba2, nsup=None, group=0, ngroups=1):
    """``tauofr`` (tauofr_utils.mod.F90:42-111) with tauadd (:139-173).  Returns tau (nlsd, nnr1): the
    group's partial sum (cp_grp_redist :101-105 is the caller's)."""
    nstate = c0.shape[0]
    nlsd = 1 if nsup is None else 2
    tau = np.zeros((nlsd, geo.nnr1))
    for is1, is2 in state_pairs(nstate, group, ngroups):
        coef1 = 0.5 * tpiba2 * f[is1] / omega                                  # :147
        coef2 = 0.0 if is2 is None else 0.5 * tpiba2 * f[is2] / omega          # :148-152
        for k in range(3):                                                     # :86-100
            psi = invfftn_sparse(geo, dpsisc(geo, gk, c0[is1], None if is2 is None else c0[is2], k))
            r1, r2 = psi.imag, psi.real                                        # :161-162
            if nsup is None:
                tau[0] += coef1 * r1 * r1 + coef2 * r2 * r2                    # :171
            else:
                sp1 = 1 if (is1 + 1) > nsup else 0                             # :156
                sp2 = 1 if ((nstate + 1) if is2 is None else (is2 + 1)) > nsup else 0   # :157
                tau[sp1] += coef1 * r1 * r1                                    # :163
                tau[sp2] += coef2 * r2 * r2                                    # :164
    return tau


def vtaupsi(geo: Geometry, c0, c2, f, gk, vtau, tpiba2, nsup=None, group=0, ngroups=1):
    """``vtaupsi`` (vtaupsi_utils.mod.F90:38-92) with taupot (:94-129) and ftauadd (:131-165).
    vtau: (ispin, nnr1).  Returns the updated c2 (a copy)."""
    nstate = c0.shape[0]
    out = c2.copy()
    vt = np.atleast_2d(vtau)
    for is1, is2 in state_pairs(nstate, group, ngroups):
        fi1 = 0.25 * f[is1] * tpiba2
        fi2 = 0.0 if is2 is None else 0.25 * f[is2] * tpiba2
        for k in range(3):
            psi = invfftn_sparse(geo, dpsisc(geo, gk, c0[is1], None if is2 is None else c0[is2], k))
            if nsup is None:
                psi = psi * vt[0]                                              # :104-107
            else:
                i2 = (nstate + 1) if is2 is None else (is2 + 1)
                if i2 <= nsup:
                    psi = psi * vt[0]                                          # :109-113
                elif (is1 + 1) > nsup:
                    psi = psi * vt[1]                                          # :114-118
                else:
                    psi = psi.real * vt[1] + 1j * psi.imag * vt[0]             # :119-125
            psi = fwfftn_sparse(geo, psi)
            psin = psi[geo.nzhs - 1]
            psii = psi[geo.indzs - 1]
            fp = psin + psii
            fm = psin - psii
            g = gk[:, k]
            out[is1] -= fi1 * g * (fm.real + 1j * fp.imag)                     # :147-148 / :160-161
            if is2 is not None:
                out[is2] -= fi2 * g * (fm.imag - 1j * fp.real)                 # :162-163
    return out


def e_test(geo: Geometry, rho_out, vpot, omega):
    """The synthetic "total energy" used for the 1e-9 Ha criterion:
    E_test = ekin + (Omega/N) * sum_r V(r) rho(r)."""
    n1, n2, n3 = geo.nr
    return rho_out["ekin"] + omega / float(n1 * n2 * n3) * float(np.dot(vpot, rho_out["rhoe"]))


# ----------------------------------------------------------------------------------------------
# synthetic inputs — oracle-side copy, cross-checked against cpmd_b200.synthetic
# ----------------------------------------------------------------------------------------------

def synthetic_inputs(geo: Geometry, nstate, seed=None, f_pattern="all2"):
    n1, n2, n3 = geo.nr
    n = n1
    if seed is None:
        seed = 1234 + n + 7 * nstate
    rng = np.random.default_rng(seed)
    gcutw = (min(geo.nr) / 4.0) ** 2
    damp = np.exp(-geo.hg / (0.25 * gcutw))
    c0 = np.empty((nstate, geo.ngw), dtype=np.complex128)
    for i in range(nstate):
        re = rng.standard_normal(geo.ngw)
        im = rng.standard_normal(geo.ngw)
        c = (re + 1j * im) * damp
        if geo.geq0:
            c[0] = c[0].real
        c /= np.sqrt(dotp(geo, c, c))
        c0[i] = c
    f = np.full(nstate, 2.0)
    if f_pattern == "mixed":
        f[::3] = 1.0
        f[1::5] = 0.0
    v = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]))
    v[:n3, :n2, :n1] = -rng.random((n3, n2, n1))
    return c0, f, v.reshape(-1)


# ----------------------------------------------------------------------------------------------
# Hartree-Fock exchange: hfx_old with func1%mhfx = 1, Gamma point, no LSD, no
# Wannier / integral screening, one task, one group (hfx_utils.mod.F90:80-965)
# ----------------------------------------------------------------------------------------------

def _psi_real(geo_w: Geometry, c):
    """psia = 0; set_psi_1_state_g(zone, c, psia); invfftn(psia,.TRUE.) (hfx_utils.mod.F90:490-491): the
    real-space state on the wavefunction FFT set, box-centre convention (no phasen), padded, complex."""
    return invfftn_sparse(geo_w, set_psi_1_state_g(geo_w, c))


def _hfx_pair(geo_w, geo_d, psia, psib, iran, pf, scgx, omega):
    """hfxab (hfx_utils.mod.F90:1034-1110): returns (ehfx, dc2a, dc2b) with c2a += dc2a, c2b += dc2b.
    ``psib`` carries the partner state in its real (iran = 1) or imaginary (iran = 2) part."""
    b = psib.real if iran == 1 else psib.imag
    psic = (psia.real * b).astype(np.complex128) / omega                       # :1052-1063
    g = fwfftn_dense(geo_d, psic)                                              # :1064
    rg = g[geo_d.nzhs - 1]
    vpotg = -pf * scgx * rg                                                    # :1068
    ehfx = float(np.sum(4.0 * vpotg * np.conj(rg)).real)                       # :1069
    if geo_d.geq0:
        ehfx -= float((2.0 * vpotg[0] * np.conj(rg[0])).real)                  # :1071
    vpotr = g_to_r(geo_d, vpotg).real                                          # :1072-1084
    psic = vpotr * (psia.real + 1j * b)                                        # :1085-1095
    r = fwfftn_sparse(geo_w, psic)                                             # :1096
    fp = r[geo_w.nzhs - 1] + r[geo_w.indzs - 1]
    fm = r[geo_w.nzhs - 1] - r[geo_w.indzs - 1]
    dc2b = -(fp.real + 1j * fm.imag)                                           # :1103
    dc2a = -(fp.imag - 1j * fm.real)                                           # :1104
    return ehfx, dc2a, dc2b


def _hfx_diag(geo_w, geo_d, psia, pf, scgx, omega):
    """hfxaa (hfx_utils.mod.F90:1203-1260): returns (ehfx, dc2a)."""
    psic = (psia.real * psia.real).astype(np.complex128) / omega               # :1218-1222
    g = fwfftn_dense(geo_d, psic)
    rg = g[geo_d.nzhs - 1]
    vpotg = -pf * scgx * rg                                                    # :1228
    ehfx = float(np.sum(2.0 * vpotg * np.conj(rg)).real)                       # :1229
    if geo_d.geq0:
        ehfx -= float((vpotg[0] * np.conj(rg[0])).real)                        # :1231
    vpotr = g_to_r(geo_d, vpotg).real
    r = fwfftn_sparse(geo_w, vpotr * psia.real.astype(np.complex128))          # :1245-1249
    fp = r[geo_w.nzhs - 1] + r[geo_w.indzs - 1]
    fm = r[geo_w.nzhs - 1] - r[geo_w.indzs - 1]
    return ehfx, -(fp.real + 1j * fm.imag)                                     # :1256


def hfx(geo_w: Geometry, geo_d: Geometry, c0, c2, f, scgx, omega, pfl=0.25):
    """``hfx_old(c0,c2,f,psia,nstate,ehfx,vhfx)`` (hfx_utils.mod.F90:80-965) for func1%mhfx = 1, Gamma point,
    cntl%tlsd = .FALSE. (pfl = 0.25, times func3%phfx for a hybrid: pass it in ``pfl``), hfxc3%twscr =
    .FALSE., one task and one group: every occupied state ia (f >= 1e-6, :474) gets its diagonal term hfxaa
    with pfx = pfl f(ia)^2 (:497-502) and every unordered pair (ia, ib) of occupied states one hfxab with pfx
    = pfl f(ia) f(ib) (:709-752; part_1d_symm_holds_pair picks one of the two orders - the result does not
    depend on which, nor on the packing of two partners into one transform, hfxab2).  ``geo_w``: the
    wavefunction FFT set (nzfs/inzs), ``geo_d``: the set of the pair densities (nzff/inzf, jhg = geo_d.ngw
    vectors) with the Coulomb kernel ``scgx``.  Returns (c2_new, ehfx, vhfx): c2 += C2_hfx (:819), ehfx =
    omega * sum (:905), vhfx = sum_ia dotp(c0_ia, c2_new_ia) (:907-909)."""
    nstate = c0.shape[0]
    occ = [i for i in range(nstate) if f[i] >= 1.0e-6]
    psi = {i: _psi_real(geo_w, c0[i, :geo_w.ngw]) for i in occ}
    c2_hfx = np.zeros((nstate, geo_w.ngw), dtype=np.complex128)
    ehfx = 0.0
    for ia in occ:
        e, d = _hfx_diag(geo_w, geo_d, psi[ia], pfl * f[ia] * f[ia], scgx, omega)
        ehfx += e
        c2_hfx[ia] += d
        for ib in occ:
            if ib <= ia:
                continue
            e, da, db = _hfx_pair(geo_w, geo_d, psi[ia], psi[ib], 1, pfl * f[ia] * f[ib], scgx, omega)
            ehfx += e
            c2_hfx[ia] += da
            c2_hfx[ib] += db
    out = np.array(c2, dtype=np.complex128, copy=True)
    out[:, :geo_w.ngw] += c2_hfx
    ehfx *= omega
    vhfx = sum(dotp(geo_w, c0[i, :geo_w.ngw], out[i, :geo_w.ngw]) for i in range(nstate))
    return out, ehfx, vhfx


def hfx_coulomb_kernel(geo_d: Geometry, tpiba2):
    """A Coulomb kernel for synthetic HFX inputs: scgx = 4 pi / (tpiba2 hg), 0 at G = 0 (the reference
    builds scgx in hfx_drivers / cppt with its own G = 0 treatment; the library takes it as an input)."""
    s = np.zeros(geo_d.ngw)
    nz = geo_d.hg > 1.0e-12
    s[nz] = 4.0 * np.pi / (tpiba2 * geo_d.hg[nz])
    return s
