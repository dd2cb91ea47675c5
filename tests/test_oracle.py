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
