"""Known-answer tests that pin the oracle (the reference ships no fixtures for this path:
SURVEY 4 / 8c).  Each test is derived from a cited line of the reference's arithmetic."""
import numpy as np
import pytest

from oracle import cpmd_oracle as orc
from oracle import staged


@pytest.mark.parametrize("n,ngw,nrays,band", [(64, 8536, 793, 31), (72, 12152, 1005, 35),
                                               (120, 56466, 2809, 59), (192, 231376, 7209, 95)])
def test_geometry_counts(n, ngw, nrays, band):
    # numbers computed in SURVEY 8 from loadpa_utils.mod.F90:286-335 / fftprp_utils.mod.F90:145-192
    g = orc.make_geometry(n)
    assert (g.ngw, g.nrays, g.kr3max - g.kr3min + 1) == (ngw, nrays, band)
    assert g.geq0 and g.kr == (n + 1,) * 3 and g.nnr1 == (n + 1) ** 3


def test_maps_are_consistent():
    g = orc.make_geometry(20)
    # +G and -G never collide except for G=0 (state_utils.mod.F90:184-187)
    assert g.nzhs[0] == g.indzs[0]
    allidx = np.concatenate([g.nzhs, g.indzs[1:]])
    assert len(np.unique(allidx)) == len(allidx)
    # ray storage index decodes back to inyh(1) (fftprp_utils.mod.F90:278)
    assert np.array_equal((g.nzhs - 1) % g.kr[0] + 1, g.inyh[0])
    assert np.array_equal((g.indzs - 1) % g.kr[0] + 1, 2 * (20 // 2 + 1) - g.inyh[0])


def _coords(g):
    n1, n2, n3 = g.nr
    z, y, x = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    return x, y, z


def test_single_plane_wave_density():
    """c0(ig0) = e^{i phi}: psi(r) = +-2 cos(G.r + phi) so rho = (f/Omega) 4 cos^2 and the charge is
    2 f = f*dotp(c,c) (state_utils.mod.F90:184-185)."""
    n, omega, f = 16, 2.5, np.array([1.3])
    g = orc.make_geometry(n)
    ig0, phi = 17, 0.4
    c0 = np.zeros((1, g.ngw), complex)
    c0[0, ig0] = np.exp(1j * phi)
    out = orc.rhoofr(g, c0, f, omega, 1.0)
    i, j, k = (g.inyh[:, ig0] - (n // 2 + 1))
    x, y, z = _coords(g)
    ref = f[0] / omega * 4.0 * np.cos(2 * np.pi * (i * x + j * y + k * z) / n + phi) ** 2
    rho = out["rhoe"].reshape(g.kr[2], g.kr[1], g.kr[0])
    assert np.abs(rho[:n, :n, :n] - ref).max() < 1e-13
    assert abs(out["rsum_r"] - 2 * f[0]) < 1e-12 and abs(out["rsum_g"] - 2 * f[0]) < 1e-12
    # pads stay exactly zero
    assert not rho[n:].any() and not rho[:, n:].any() and not rho[:, :, n:].any()


def test_g0_only_density():
    g = orc.make_geometry(16)
    c0 = np.zeros((1, g.ngw), complex)
    c0[0, 0] = 1.0
    out = orc.rhoofr(g, c0, np.array([2.0]), 4.0, 1.0)
    rho = out["rhoe"].reshape(17, 17, 17)[:16, :16, :16]
    assert np.abs(rho - 0.5).max() < 1e-14     # f/Omega, entry written once (state_utils :187)


@pytest.mark.parametrize("n,nstate", [(16, 4), (20, 5), (24, 3)])
def test_charge_and_energy_identities(n, nstate):
    g = orc.make_geometry(n)
    c0, f, v = orc.synthetic_inputs(g, nstate)
    omega, tpiba2 = 3.0, 0.7
    out = orc.rhoofr(g, c0, f, omega, tpiba2)
    # rhoofr_utils.mod.F90:607-619 (reference tolerance 1e-6)
    assert abs(out["rsum_r"] - out["rsum_g"]) < 1e-12
    c2 = orc.vpsi(g, c0, np.zeros_like(c0), f, v, tpiba2)
    lhs = -sum(orc.dotp(g, c0[i], c2[i]) for i in range(nstate))
    assert abs(lhs - orc.e_test(g, out, v, omega)) < 1e-11


def test_constant_potential():
    """V == v0  =>  c2_i = -f_i (tpiba2 hg / 2 + v0) c0_i  (vpsi_utils.mod.F90:666-670, fi=f/2)."""
    g = orc.make_geometry(16)
    c0, f, _ = orc.synthetic_inputs(g, 3)
    f = np.array([2.0, 1.0, 0.5])
    v0, tpiba2 = -0.37, 1.9
    v = np.zeros((17, 17, 17))
    v[:16, :16, :16] = v0
    c2 = orc.vpsi(g, c0, np.zeros_like(c0), f, v.reshape(-1), tpiba2)
    ref = -f[:, None] * (0.5 * tpiba2 * g.hg[None, :] + v0) * c0
    assert np.abs(c2 - ref).max() < 1e-13


def test_pair_packing_invariance_and_odd_block():
    g = orc.make_geometry(16)
    c0, f, v = orc.synthetic_inputs(g, 3)           # odd: last state takes set_psi_1_state_g
    full = orc.vpsi(g, c0, np.zeros_like(c0), f, v, 1.0)
    rho_full = orc.rhoofr(g, c0, f, 1.0, 1.0)["rhoe"]
    rho_sum = np.zeros_like(rho_full)
    for i in range(3):
        one = orc.vpsi(g, c0[i:i + 1], np.zeros_like(c0[i:i + 1]), f[i:i + 1], v, 1.0)
        assert np.abs(one[0] - full[i]).max() < 1e-13
        rho_sum += orc.rhoofr(g, c0[i:i + 1], f[i:i + 1], 1.0, 1.0)["rhoe"]
    assert np.abs(rho_sum - rho_full).max() < 1e-13


def test_zero_occupation_semantics():
    g = orc.make_geometry(16)
    c0, f, v = orc.synthetic_inputs(g, 4)
    f = np.array([0.0, 0.0, 2.0, 0.0])
    out = orc.rhoofr(g, c0, f, 1.0, 1.0)            # pair (0,1) skipped, pair (2,3) computed
    only = orc.rhoofr(g, c0[2:3], f[2:3], 1.0, 1.0)
    assert np.abs(out["rhoe"] - only["rhoe"]).max() < 1e-13
    # vpsi still transforms f=0 states and substitutes fi=1 (0.5 with tksham): :627-633
    c2 = orc.vpsi(g, c0, np.zeros_like(c0), f, v, 1.0)
    c2k = orc.vpsi(g, c0, np.zeros_like(c0), f, v, 1.0, tksham=True)
    assert np.abs(c2[0]).max() > 0 and np.abs(c2[0] - 2.0 * c2k[0]).max() < 1e-13
    assert np.abs(c2[2] - c2k[2]).max() == 0.0


def test_part_1d_blocks_cover_all_states():
    for n in (1, 7, 64, 130):
        for ng in (1, 2, 3, 8):
            seen = []
            for grp in range(ng):
                cnt = orc.part_1d_nbr_el_in_blk(n, grp, ng)
                seen += [orc.part_1d_get_el_in_blk(i, n, grp, ng) for i in range(1, cnt + 1)]
            assert seen == list(range(1, n + 1))


def test_group_partition_sums_to_full():
    g = orc.make_geometry(16)
    c0, f, v = orc.synthetic_inputs(g, 7, f_pattern="mixed")
    full = orc.rhoofr(g, c0, f, 1.0, 1.0)["rhoe"]
    c2_full = orc.vpsi(g, c0, np.zeros_like(c0), f, v, 1.0)
    for ng in (2, 3):
        acc = np.zeros_like(full)
        c2 = np.zeros_like(c0)
        for grp in range(ng):
            acc += orc.rhoofr(g, c0, f, 1.0, 1.0, grp, ng)["rhoe"]
            c2 = orc.vpsi(g, c0, c2, f, v, 1.0, grp, ng)
        assert np.abs(acc - full).max() < 1e-13
        assert np.abs(c2 - c2_full).max() < 1e-13


@pytest.mark.parametrize("n,nstate", [(16, 4), (20, 5), (30, 3), (36, 4), (84, 3)])
def test_staged_c_restatement_matches_dense(n, nstate):
    """fftnew's staged sparse pipeline (C) == dense 3-D FFT (NumPy) on the same inputs."""
    g = orc.make_geometry(n)
    c0, f, v = orc.synthetic_inputs(g, nstate, f_pattern="mixed" if nstate > 4 else "all2")
    a = orc.rhoofr(g, c0, f, 1.3, 0.9)
    b = staged.rhoofr(g, c0, f, 1.3, 0.9)
    assert np.abs(a["rhoe"] - b["rhoe"]).max() <= 1e-13 * np.abs(a["rhoe"]).max()
    for k in ("ekin", "rsum_g", "rsum_r"):
        assert abs(a[k] - b[k]) < 1e-11 * max(1.0, abs(a[k]))
    c2 = 0.3 * c0
    ca = orc.vpsi(g, c0, c2, f, v, 0.9)
    cb = staged.vpsi(g, c0, c2, f, v, 0.9)
    assert np.abs(ca - cb).max() <= 1e-13 * np.abs(ca).max()


def test_anisotropic_mesh():
    g = orc.make_geometry((16, 20, 24))
    c0, f, v = orc.synthetic_inputs(g, 3)
    a = orc.rhoofr(g, c0, f, 1.0, 1.0)
    b = staged.rhoofr(g, c0, f, 1.0, 1.0)
    assert abs(a["rsum_r"] - a["rsum_g"]) < 1e-12
    assert np.abs(a["rhoe"] - b["rhoe"]).max() <= 1e-13 * np.abs(a["rhoe"]).max()


def test_lsd_known_answers():
    """LSD restatement (rhoofr_utils.mod.F90:375-385,543-559; vpsi_utils.mod.F90:450-482) pinned by
    identities that follow from the reference formulas: the total density does not depend on the
    spin labels; nsup = nstate / 0 put everything in one channel; csums is the alpha-beta charge;
    each state's C2 is the non-LSD result with the potential of its own spin."""
    geo = orc.make_geometry(16)
    ns = 7
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    ref = orc.rhoofr(geo, c0, f, 1.0, 1.0)
    occ = np.array([f[i] * orc.dotp(geo, c0[i], c0[i]) for i in range(ns)])
    for nsup in (0, 3, 4, 7):
        o = orc.rhoofr_lsd(geo, c0, f, 1.0, 1.0, nsup)
        assert np.abs(o["rhoe"][0] - ref["rhoe"]).max() < 1e-13 * np.abs(ref["rhoe"]).max()
        assert abs(o["rsum_r"] - ref["rsum_r"]) < 1e-12 and o["ekin"] == ref["ekin"]
        assert abs(o["csums"] - (occ[:nsup].sum() - occ[nsup:].sum())) < 1e-11
        assert o["csumsabs"] >= abs(o["csums"]) - 1e-12 and (o["rhoe"][1] >= 0).all()
        if nsup == ns:
            assert not o["rhoe"][1].any()
    v2 = np.stack([v, 0.5 * v[::-1]])
    a = orc.vpsi(geo, c0, np.zeros_like(c0), f, v2[0], 1.0)
    b = orc.vpsi(geo, c0, np.zeros_like(c0), f, v2[1], 1.0)
    for nsup in (0, 3, 4, 7):
        o = orc.vpsi_lsd(geo, c0, np.zeros_like(c0), f, v2, 1.0, nsup)
        exp = np.where((np.arange(ns) < nsup)[:, None], a, b)
        assert np.abs(o - exp).max() < 1e-13 * np.abs(exp).max()
    # groups: partial channels add up, finish after the sum (cp_grp_redist before :543)
    acc = np.zeros((2, geo.nnr1))
    for g in range(3):
        acc += orc.rhoofr_lsd(geo, c0, f, 1.0, 1.0, 3, group=g, ngroups=3)["rhoe"]
    full = orc.rhoofr_lsd(geo, c0, f, 1.0, 1.0, 3)
    rr, cs, ca = orc.lsd_finish(geo, acc, 1.0)
    assert np.abs(acc - full["rhoe"]).max() < 1e-13 * np.abs(full["rhoe"]).max()
    assert abs(cs - full["csums"]) < 1e-12 and abs(ca - full["csumsabs"]) < 1e-12


# ---------------------------------------------------------------------------------------------
# dense transforms + local part of vofrho: known answers from the cited formulas
# ---------------------------------------------------------------------------------------------

def _mesh_field(geo, fn):
    n1, n2, n3 = geo.nr
    z, y, x = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    a = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]))
    a[:n3, :n2, :n1] = fn(x, y, z)
    return a.reshape(-1)


def _find_g(geo, g):
    nh = [n // 2 + 1 for n in geo.nr]
    m = (geo.inyh[0] == nh[0] + g[0]) & (geo.inyh[1] == nh[1] + g[1]) & (geo.inyh[2] == nh[2] + g[2])
    idx = np.nonzero(m)[0]
    return int(idx[0]) if idx.size else None


def test_dense_transform_known_answers():
    """phasen + dense fwfftn put G at inyh = nh + G (fftmain_utils.mod.F90:137-153,
    fftutil_utils.mod.F90:479-503): a cosine has exactly two coefficients 1/2 (one in the half
    sphere), a constant only G=0, and the inverse reproduces the field with zero imaginary part."""
    n = 16
    geo = orc.make_density_geometry(n)
    assert geo.geq0 and geo.kr3min == 2 and geo.kr3max == n      # |g| <= n/2 - 1 on every axis
    wgeo = orc.make_geometry(n)
    assert np.array_equal(geo.inyh[:, :wgeo.ngw], wgeo.inyh)     # first ngw of nhg = wavefunction sphere
    g = (2, -3, 1)
    rho = _mesh_field(geo, lambda x, y, z: 0.7 + np.cos(2 * np.pi * (g[0] * x + g[1] * y + g[2] * z) / n + 0.4))
    rg = orc.rho_to_g(geo, rho)
    ig = _find_g(geo, g)
    want = np.zeros(geo.ngw, complex)
    want[0] = 0.7
    want[ig] = 0.5 * np.exp(0.4j)
    assert np.abs(rg - want).max() < 1e-15
    back = orc.g_to_r(geo, rg)
    assert np.abs(back.real - rho).max() < 1e-14 and np.abs(back.imag).max() < 1e-14
    # pads stay zero
    b3 = back.reshape(geo.kr[2], geo.kr[1], geo.kr[0])
    assert not b3[n:].any() and not b3[:, n:].any() and not b3[:, :, n:].any()


def test_density_from_rhoofr_has_nel_at_g0():
    """rhog(G=0) * omega = number of electrons: links rhoofr's charge check
    (rhoofr_utils.mod.F90:607-619) with the dense forward transform's 1/N scale."""
    n, ns, omega = 16, 4, 1.7
    geo = orc.make_geometry(n)
    c0, f, _ = orc.synthetic_inputs(geo, ns)
    r = orc.rhoofr(geo, c0, f, omega, 1.0)
    dgeo = orc.make_density_geometry(n)
    rg = orc.rho_to_g(dgeo, r["rhoe"])
    assert abs(rg[0].real * omega - r["rsum_g"]) < 1e-12 and abs(rg[0].imag) < 1e-15
    # |psi|^2 of functions band limited to |G| < n/4 is band limited to |G| < n/2: the density
    # sphere holds all of it except the corners of the cube outside the sphere -> round trip is
    # close but not exact; the part inside the sphere is reproduced exactly
    back = orc.g_to_r(dgeo, rg).real
    assert np.abs(orc.rho_to_g(dgeo, back) - rg).max() < 1e-15


def test_hartree_known_answer_and_energies():
    """ppener (ppener_utils.mod.F90:85-104) with eirop = eivps = 0 and scg = 4 pi/(tpiba2 G^2): the
    potential of rho = a cos(G.r) is scg(G) * rho, eh = sum scg |rho_G|^2 over the half sphere =
    (1/2N) sum_r rho(r) V(r), ee = eh, ei = eps = 0."""
    n = 20
    geo = orc.make_density_geometry(n)
    g = (1, 2, -2)
    a = 0.3
    rho = _mesh_field(geo, lambda x, y, z: a * np.cos(2 * np.pi * (g[0] * x + g[1] * y + g[2] * z) / n))
    scg, _, _ = orc.synthetic_vofrho_inputs(geo, tpiba2=0.9)
    zero = np.zeros(geo.ngw, complex)
    r = orc.vofrho_local(geo, rho, scg, zero, zero)
    k = 4 * np.pi / (0.9 * 9.0)
    assert np.abs(r["v"] - k * rho).max() < 1e-14
    assert abs(r["eh"] - k * (a / 2) ** 2) < 1e-15 and abs(r["ee"] - r["eh"]) < 1e-18
    assert r["ei"] == 0 and r["eps"] == 0 and r["vploc"] == 0
    nn = float(n ** 3)
    assert abs(r["eh"].real - 0.5 * np.dot(rho, r["v"]) / nn) < 1e-15


def test_ppener_g0_special_case_and_epseu():
    """G=0 entry (ppener_utils.mod.F90:58-70): half weights, vtemp(1) = scg(1)*rhog without vps;
    eps = sum conj(rho_G) vps_G over the half sphere = (1/2N) sum_r rho(r) Vps(r) for real fields,
    so that epseu = 2 Re(eps) omega is the integral of rho * Vps (vofrhoa_utils.mod.F90:127)."""
    n = 16
    geo = orc.make_density_geometry(n)
    rng = np.random.default_rng(5)
    scg, eivps, eirop = orc.synthetic_vofrho_inputs(geo, seed=9)
    scg = scg.copy()
    scg[0] = 0.37                                            # exercise the G=0 formulas
    rhog = orc.rho_to_g(geo, orc.g_to_r(geo, (rng.standard_normal(geo.ngw) + 1j * rng.standard_normal(geo.ngw)
                                              ) * np.exp(-geo.hg / 20)).real)
    eh, ei, ee, eps, vploc, vtemp = orc.ppener(geo, rhog, scg, eivps, eirop)
    assert vploc == eivps[0].real
    assert vtemp[0] == scg[0] * (rhog[0] + eirop[0])
    assert np.abs(vtemp[1:] - (scg[1:] * (rhog[1:] + eirop[1:]) + eivps[1:])).max() == 0
    # real-space cross-check of eps: Vps(r) from the same scatter/inverse as the potential
    rho_r = orc.g_to_r(geo, rhog).real
    vps_r = orc.g_to_r(geo, eivps).real
    nn = float(n ** 3)
    assert abs(eps.real - 0.5 * np.dot(rho_r, vps_r) / nn) < 1e-13
    # Hartree energy of the total (electron + Gaussian ion) charge, with the G=0 half weight
    tot = rhog + eirop
    want = 0.5 * scg[0] * tot[0].real ** 2 + np.sum(scg[1:] * np.abs(tot[1:]) ** 2)
    assert abs(eh - want) < 1e-13


def test_vofrho_golden_vectors_frozen():
    from helpers import golden_vofrho_cases, load_golden_vofrho
    cases = golden_vofrho_cases()
    assert cases
    for path in cases:
        d = load_golden_vofrho(path)
        geo = orc.fft_maps(d["nr"], d["inyh"], d["hg"])
        assert np.array_equal(geo.nzhs, d["nzh"]) and np.array_equal(geo.indzs, d["indz"])
        r = orc.vofrho_local(geo, d["rhoe"], d["scg"], d["eivps"], d["eirop"])
        assert np.abs(r["v"] - d["v"]).max() <= 1e-13 * np.abs(d["v"]).max()
        assert np.abs(r["rhog"] - d["rhog"]).max() <= 1e-14


# ---------------------------------------------------------------------------------------------
# k-points: known answers
# ---------------------------------------------------------------------------------------------

def test_kpt_reduces_to_gamma_at_k0():
    """A Gamma state c written as the k-point state [c, conj(c)] with k = 0 (hgkp = hgkm = hg) must
    give the single-state Gamma results: same rho, ekin, charge, and C2 = [c2, conj(c2)]
    (vpsi_utils.mod.F90:614-625 against :655-671 with one state per transform)."""
    geo = orc.make_geometry(16)
    cg, fg, vg = orc.synthetic_inputs(geo, 3, f_pattern="mixed")
    fg = fg.copy()
    ck = np.concatenate([cg, np.conj(cg)], axis=1)
    ck[:, geo.ngw] = 0.0
    rk = orc.rhoofr_kpt(geo, ck, fg, 1.0, geo.hg, geo.hg, 1.7, 0.8)
    rg = orc.rhoofr(geo, cg, fg, 1.7, 0.8)
    assert np.abs(rk["rhoe"] - rg["rhoe"]).max() < 1e-13 * np.abs(rg["rhoe"]).max()
    assert abs(rk["ekin"] - rg["ekin"]) < 1e-12 and abs(rk["rsum_g"] - rg["rsum_g"]) < 1e-12
    # f = 0: Gamma substitutes fi = 1 for f/2, the k-branch fi = 2 for f: the same force
    occ = np.ones_like(fg, dtype=bool)
    assert (fg == 0).any()
    c2k = orc.vpsi_kpt(geo, ck, np.zeros_like(ck), fg, geo.hg, geo.hg, vg, 0.8)
    c2g = orc.vpsi(geo, cg, np.zeros_like(cg), fg, vg, 0.8)
    assert np.abs(c2k[occ, :geo.ngw] - c2g[occ]).max() < 1e-14
    assert np.abs(c2k[occ, geo.ngw + 1:] - np.conj(c2g[occ, 1:])).max() < 1e-14
    assert not c2k[:, geo.ngw].any()                                  # :625


def test_kpt_known_answers():
    """Single plane wave at -G: psi(r) = exp(-iG.r) (up to the centre-origin sign), so rho = wk f/omega
    everywhere; constant potential: C2 = -f (tpiba2/2 |k+-G|^2 + v0) c0; charge identity."""
    n = 16
    geo = orc.make_geometry(n)
    ngw = geo.ngw
    c0 = np.zeros((1, 2 * ngw), complex)
    c0[0, ngw + 7] = np.exp(0.3j)
    _, _, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, 1)
    f = np.array([2.0])
    r = orc.rhoofr_kpt(geo, c0, f, 0.5, hgkp, hgkm, 1.5, 1.0)
    box = r["rhoe"].reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n, :n, :n]
    assert np.abs(box - 0.5 * 2.0 / 1.5).max() < 1e-14
    assert abs(r["ekin"] - 0.5 * 0.5 * 2.0 * hgkm[7]) < 1e-14
    c0r, f5, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, 5)
    r = orc.rhoofr_kpt(geo, c0r, f5, 0.7, hgkp, hgkm, 1.5, 1.0)
    assert abs(r["rhoe"].sum() * 1.5 / n ** 3 - r["rsum_g"]) < 1e-12
    v0 = np.zeros_like(v)
    v0.reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n, :n, :n] = -0.37
    c2 = orc.vpsi_kpt(geo, c0r, np.zeros_like(c0r), f5, hgkp, hgkm, v0, 0.9)
    fi = np.where(f5 == 0, 2.0, f5)[:, None]
    want = -fi * (0.5 * 0.9 * np.concatenate([hgkp, hgkm]) - 0.37) * c0r
    want[:, ngw] = 0.0
    assert np.abs(c2 - want).max() < 1e-14
    # energy identity per k-point: -Re sum conj(c0) c2 = 2 ekin/wk-part + int V rho (occupied states, fi = f)
    occ = f5 != 0
    c2 = orc.vpsi_kpt(geo, c0r, np.zeros_like(c0r), f5, hgkp, hgkm, v, 1.0)
    r1 = orc.rhoofr_kpt(geo, c0r, f5, 1.0, hgkp, hgkm, 1.0, 1.0)
    lhs = -np.sum((np.conj(c0r[occ]) * c2[occ]).real)
    rhs = r1["ekin"] + np.dot(v, r1["rhoe"]) / n ** 3
    assert abs(lhs - rhs) < 1e-12


# ---------------------------------------------------------------------------------------------
# meta-GGA tauofr / vtaupsi: known answers
# ---------------------------------------------------------------------------------------------

def test_tau_known_answers():
    """int tau(r) dr = ekin (tauadd's tpiba2 f / (2 omega) |grad psi|^2 against kin_energy); a single
    plane wave gives tau = (f tpiba2 / 2 omega) 4 |G|^2 sin^2(G.r + phi); constant vtau = v0 turns vtaupsi
    into c2 -= f/2 v0 tpiba2 |G|^2 c0; and -sum dotp(c0, dC2) = (1/N) sum vtau(r) tau(r) omega."""
    n = 16
    geo = orc.make_geometry(n)
    gk = orc.gk_cartesian(geo)
    assert np.allclose((gk ** 2).sum(axis=1), geo.hg) and not gk[0].any()
    c0, f, v = orc.synthetic_inputs(geo, 3, f_pattern="mixed")
    f[1] = 2.0
    omega, tp = 1.7, 0.8
    tau = orc.tauofr(geo, c0, f, gk, omega, tp)[0]
    r = orc.rhoofr(geo, c0, f, omega, tp)
    assert abs(tau.sum() * omega / n ** 3 - r["ekin"]) < 1e-12
    assert tau.min() >= 0.0
    # single plane wave
    ig, phi = 9, 0.4
    c1 = np.zeros((1, geo.ngw), complex)
    c1[0, ig] = np.exp(1j * phi)
    t1 = orc.tauofr(geo, c1, np.array([2.0]), gk, omega, tp)[0].reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n, :n, :n]
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    arg = 2 * np.pi * (gk[ig, 0] * x + gk[ig, 1] * y + gk[ig, 2] * z) / n + phi
    want = 2.0 * tp / (2 * omega) * 4 * geo.hg[ig] * np.sin(arg) ** 2
    assert np.abs(t1 - want).max() < 1e-13
    # constant potential
    v0 = np.zeros_like(v)
    v0.reshape(geo.kr[2], geo.kr[1], geo.kr[0])[:n, :n, :n] = -0.37
    d = orc.vtaupsi(geo, c0, np.zeros_like(c0), f, gk, v0[None], tp)
    assert np.abs(d + f[:, None] * 0.5 * (-0.37) * tp * geo.hg * c0).max() < 1e-15
    # energy identity with a general potential
    d = orc.vtaupsi(geo, c0, np.zeros_like(c0), f, gk, v[None], tp)
    lhs = -sum(orc.dotp(geo, c0[i], d[i]) for i in range(3))
    assert abs(lhs - np.dot(v, tau) * omega / n ** 3) < 1e-12
    # LSD: the channels add up to the unpolarised tau, pair packing is irrelevant
    for nsup in (0, 1, 2, 3):
        t2 = orc.tauofr(geo, c0, f, gk, omega, tp, nsup)
        assert np.abs(t2.sum(axis=0) - tau).max() < 1e-14
        d2 = orc.vtaupsi(geo, c0, np.zeros_like(c0), f, gk, np.stack([v, v]), tp, nsup)
        assert np.abs(d2 - d).max() < 1e-15


@pytest.mark.parametrize("nr,ns", [(16, 5), ((16, 20, 24), 4), (30, 7)])
def test_pocketfft_staged_restatement_matches_dense(nr, ns):
    """oracle/staged_pocketfft.py (the stronger CPU baseline of bench.py: fftnew's staged sparse passes with
    a library 1-D FFT) against the dense NumPy restatement, all groupings."""
    from oracle import staged_pocketfft as spf
    geo = orc.make_geometry(nr)
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    for g, ng in ((0, 1), (0, 2), (2, 3)):
        a = orc.rhoofr(geo, c0, f, 1.3, 0.9, g, ng)
        b = spf.rhoofr(geo, c0, f, 1.3, 0.9, g, ng, batch=2)
        assert np.abs(a["rhoe"] - b["rhoe"]).max() <= 1e-13 * max(np.abs(a["rhoe"]).max(), 1e-300)
        for k in ("ekin", "rsum_g", "rsum_r"):
            assert abs(a[k] - b[k]) < 1e-11
        c2a = orc.vpsi(geo, c0, 0.5 * c0, f, v, 0.9, g, ng)
        c2b = spf.vpsi(geo, c0, 0.5 * c0, f, v, 0.9, g, ng, batch=3)
        assert np.abs(c2a - c2b).max() <= 1e-13 * np.abs(c2a).max()


def test_hfx_known_answers():
    """Hartree-Fock exchange restatement (hfx_utils.mod.F90:80-965, 1034-1260) against closed forms.
    One state, one plane wave c(G1) = 1/sqrt(2): psi(r) = +-sqrt(2) cos(G1 r), pair density psi^2 / omega =
    (1 + cos(2 G1 r)) / omega, i.e. rho(0) = 1/omega, rho(+-2 G1) = 1/(2 omega); with pf = pfl f^2
    ehfx = -omega pf [scgx(0) rho(0)^2 + 2 scgx(2 G1) rho(2 G1)^2] (hfxaa :1226-1231, times omega :905).
    Then: Euler identity sum_i dotp(c0_i, dC2_i) = -2 ehfx, degree-4 homogeneity, invariance under the order of
    the states, and additivity - unoccupied states (f < 1e-6) take no part (:474)."""
    nr = 16
    gw = orc.make_geometry(nr)
    gd = orc.make_density_geometry(nr)
    om, tp, fa = 1.7, 0.8, 2.0
    scgx = orc.hfx_coulomb_kernel(gd, tp) + 0.3          # a non-zero G = 0 entry exercises its special weight
    ig1 = int(np.nonzero((gw.inyh[0] - 9 == 1) & (gw.inyh[1] - 9 == 2) & (gw.inyh[2] - 9 == 0))[0][0])
    c0 = np.zeros((1, gw.ngw), complex)
    c0[0, ig1] = 1.0 / np.sqrt(2.0)
    out, e, v = orc.hfx(gw, gd, c0, np.zeros_like(c0), np.array([fa]), scgx, om)
    g2 = 2 * (gw.inyh[:, ig1] - 9)
    j = int(np.nonzero(np.all(gd.inyh - 9 == g2[:, None], axis=0) | np.all(gd.inyh - 9 == -g2[:, None], axis=0))[0][0])
    pf = 0.25 * fa * fa
    want = -om * pf * (scgx[0] / om ** 2 + 2.0 * scgx[j] * (0.5 / om) ** 2)
    assert abs(e - want) < 1e-13 * abs(want)
    assert abs(v + 2.0 * e) < 1e-12 * abs(e)
    c0, f, _ = orc.synthetic_inputs(gw, 5, f_pattern="mixed")
    z = np.zeros_like(c0)
    o1, e1, v1 = orc.hfx(gw, gd, c0, z, f, scgx, om)
    assert abs(v1 + 2.0 * e1) < 1e-11 * abs(e1)
    o2, e2, _ = orc.hfx(gw, gd, 1.5 * c0, z, f, scgx, om)
    assert abs(e2 - 1.5 ** 4 * e1) < 1e-11 * abs(e2) and np.abs(o2 - 1.5 ** 3 * o1).max() < 1e-11 * np.abs(o2).max()
    keep = f >= 1e-6
    o3, e3, _ = orc.hfx(gw, gd, c0[keep], z[keep], f[keep], scgx, om)
    assert abs(e3 - e1) < 1e-12 * abs(e1) and np.abs(o3 - o1[keep]).max() < 1e-13 and not o1[~keep].any()
