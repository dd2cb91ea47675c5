"""The oracle (and the kernels) against the REFERENCE'S OWN STATEMENTS for the parts of rhoofr / vpsi that are
plain Fortran loops: pairing (rhoofr_utils.mod.F90:306-310, vpsi_utils.mod.F90:376-383, part_1d.mod.F90:31-34,
51-53), coefficients and density accumulation (rhoofr_utils.mod.F90:369-374, density_utils.mod.F90:77-80), occupation
rules and the +-G unpack with the kinetic term (vpsi_utils.mod.F90:627-672), kin_energy / dotp
(kin_energy_utils.mod.F90:62-110, dotp_utils.mod.F90:45-52).  oracle/fsnip.py executes those line ranges from
/root/reference/src (no Fortran compiler in the image); tests/golden/fsnip/ holds what they produced
(tools/make_golden_fsnip.py), so the comparisons also run where the reference tree is absent."""
import numpy as np
import pytest

from cpmd_b200 import lib
from cpmd_b200.api import Plan
from helpers import ETOL, RTOL, golden_fsnip_cases, load_golden_fsnip, relmax
from oracle import cpmd_oracle as orc
from oracle import fsnip

needs_ref = pytest.mark.skipif(not fsnip.available(), reason="needs /root/reference/src")
IDS = lambda p: p.split("/")[-1][:-4]  # noqa: E731


def _geo(d):
    return orc.fft_maps(d["nr"], d["inyh"], d["hg"])


def test_fixtures_exist():
    assert len(golden_fsnip_cases()) >= 5


@pytest.mark.parametrize("path", golden_fsnip_cases(), ids=IDS)
def test_oracle_matches_reference_statements(path):
    d = load_golden_fsnip(path)
    geo = _geo(d)
    ns = d["c0"].shape[0]
    # pairing
    want = [(a - 1, None if b > ns else b - 1) for a, b in d["pairs_vpsi"].tolist()]
    assert orc.state_pairs(ns, d["group"], d["ngroups"]) == want
    assert np.array_equal(d["pairs_vpsi"], d["pairs_rhoofr"])
    # kin_energy / dotp (all states, like the reference computes them on every group)
    full = orc.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"])
    assert abs(full["ekin"] - d["ekin"]) < 1e-12 * max(1.0, abs(d["ekin"]))
    assert abs(full["rsum_g"] - d["rsum_g"]) < 1e-12 * max(1.0, abs(d["rsum_g"]))
    # density of the group's block
    blk = orc.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["group"], d["ngroups"])
    assert relmax(blk["rhoe"], d["rhoe"]) < 1e-13
    assert abs(blk["rsum_r"] - d["rsum_r"]) < 1e-12 * max(1.0, abs(d["rsum_r"]))
    # unpack + kinetic + occupation rules
    c2 = orc.vpsi(geo, d["c0"], d["c2_in"], d["f"], d["vpot"], d["tpiba2"], d["group"], d["ngroups"],
                  tksham=d["tksham"])
    assert relmax(c2, d["c2_out"]) < 1e-13


@pytest.mark.parametrize("path", golden_fsnip_cases(), ids=IDS)
def test_kernels_match_reference_statements(emu_cdll, path):
    """The CUDA sources (CPU simulator) through the C ABI against the same fixtures."""
    d = load_golden_fsnip(path)
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=2, _cdll=emu_cdll)
    rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"], ngroups=d["ngroups"], my_group=d["group"])
    assert relmax(rho, d["rhoe"]) < RTOL and abs(rr - d["rsum_r"]) < ETOL
    if d["ngroups"] == 1:
        assert abs(ekin - d["ekin"]) < ETOL and abs(rg - d["rsum_g"]) < ETOL
    c2 = d["c2_in"].copy()
    p.vpsi(d["c0"], c2, d["f"], d["vpot"], ngroups=d["ngroups"], my_group=d["group"],
           flags=lib.CPB_VPSI_TKSHAM if d["tksham"] else 0)
    assert relmax(c2, d["c2_out"]) < RTOL


@needs_ref
def test_translator_reads_the_cited_statements():
    """The line ranges still hold the statements the docstrings cite (a moved reference would silently pin nothing)."""
    s = fsnip.read_statements("vpsi_utils.mod.F90", 627, 672)
    assert s[0].replace(" ", "") == "fi=f(is1)*0.5_real_8"
    assert any("CMPLX(AIMAG(fp),-REAL(fm),kind=real_8)" in x for x in s) and s[-1].upper() == "ENDIF"
    assert fsnip.read_statements("density_utils.mod.F90", 77, 80)[1].replace(" ", "") == \
        "rho(l)=rho(l)+alpha_real*REAL(psi(l))**2+alpha_imag*AIMAG(psi(l))**2"
    assert fsnip.read_statements("rhoofr_utils.mod.F90", 369, 369) == ["coef3=crge%f(is1,1)/parm%omega"]
    assert fsnip.read_statements("kin_energy_utils.mod.F90", 110, 110) == ["ener_com%ekin=xkin*parm%tpiba2"]
    assert "ddot(2*n-2,a(2),1,b(2),1)" in fsnip.read_statements("dotp_utils.mod.F90", 45, 52)[-2]
    py = fsnip.translate(["IF (fi.EQ.0._real_8.AND..NOT.cntl%tksham) fi=1._real_8"])
    assert "fi==0.0 and  not cntl.tksham" in py and "fi = 1.0" in py


@needs_ref
@pytest.mark.parametrize("seed,nr,ns,fp,ngroups,group,tksham", [(3, 16, 6, "mixed", 1, 0, False),
                                                                 (4, 16, 3, "mixed", 2, 1, True),
                                                                 (5, (16, 20, 16), 4, "all2", 1, 0, False)])
def test_oracle_matches_reference_statements_live(seed, nr, ns, fp, ngroups, group, tksham):
    """Fresh inputs, the reference statements executed now."""
    from oracle import fsnip_cases as fc
    geo = orc.make_geometry(nr)
    c0, f, v = orc.synthetic_inputs(geo, ns, seed=seed, f_pattern=fp)
    ekin, rsum = fc.kin_energy(geo, c0, f, 0.9)
    ref = orc.rhoofr(geo, c0, f, 1.7, 0.9)
    assert abs(ekin - ref["ekin"]) < 1e-12 * max(1.0, abs(ekin)) and abs(rsum - ref["rsum_g"]) < 1e-12 * max(1.0, rsum)
    blk = orc.rhoofr(geo, c0, f, 1.7, 0.9, group, ngroups)
    assert relmax(blk["rhoe"], fc.rhoofr(geo, c0, f, 1.7, 0.9, group, ngroups)) < 1e-13
    c2 = orc.vpsi(geo, c0, 0.25 * c0, f, v, 0.9, group, ngroups, tksham=tksham)
    assert relmax(c2, fc.vpsi(geo, c0, 0.25 * c0, f, v, 0.9, group, ngroups, tksham)) < 1e-13


@needs_ref
def test_fixtures_are_reproducible(tmp_path):
    """tools/make_golden_fsnip.py regenerates the committed fixtures bit for bit."""
    import tools.make_golden_fsnip as mk
    old = mk.OUT
    mk.OUT = str(tmp_path)
    try:
        mk.main()
    finally:
        mk.OUT = old
    for path in golden_fsnip_cases():
        a, b = np.load(path), np.load(tmp_path / path.split("/")[-1])
        for k in ("rhoe", "c2_out", "ekin", "rsum_g"):
            assert np.array_equal(a[k], b[k]), (path, k)


@pytest.mark.parametrize("path", [p for p in golden_fsnip_cases() if "tksham" not in p], ids=IDS)
def test_staged_c_oracle_matches_reference_statements(path):
    """oracle/staged_oracle.c (the checker of the large meshes and the timed CPU baseline) against the same fixtures."""
    from oracle import staged
    d = load_golden_fsnip(path)
    geo = _geo(d)
    blk = staged.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["group"], d["ngroups"])
    assert relmax(blk["rhoe"], d["rhoe"]) < 1e-12
    c2 = staged.vpsi(geo, d["c0"], d["c2_in"], d["f"], d["vpot"], d["tpiba2"], d["group"], d["ngroups"])
    assert relmax(c2, d["c2_out"]) < 1e-12


@needs_ref
def test_lsd_post_processing_matches_reference_statements():
    """rhoofr_utils.mod.F90:548-558 (alpha+beta / beta, csums, csumsabs) against the oracle's lsd_finish."""
    from oracle.fsnip import FArr, ns
    geo = orc.make_geometry(16)
    rng = np.random.default_rng(11)
    r2 = rng.random((2, geo.nnr1)) - 0.2
    want = r2.copy()
    _, csums, csumsabs = orc.lsd_finish(geo, want, 3.3)
    rhoe = np.ascontiguousarray(r2.T)                               # rhoe(nnr1, 2), column-major
    env = dict(fpar=ns(nnr1=geo.nnr1), rhoe=FArr(rhoe), parm=ns(omega=3.3), lr1s=16, lr2s=16, lr3s=16,
               chrg=ns(csums=0.0, csumsabs=0.0))
    fsnip.run("rhoofr_utils.mod.F90", 548, 558, env)
    assert abs(env["chrg"].csums - csums) < 1e-12 and abs(env["chrg"].csumsabs - csumsabs) < 1e-12
    assert np.allclose(rhoe.T, want, rtol=0, atol=1e-15)


@needs_ref
def test_kpoint_variants_match_reference_statements():
    """One k-point of rhoofr_c (rhoofr_c_utils.mod.F90:117-178,182: charge and kinetic sums, the state loop with
    set_psi_1_state_g_kpts, state_utils.mod.F90:202-222, and build_density_sum) and the k-point unpack of vpsi
    (vpsi_utils.mod.F90:562-625), executed from the reference's statements, against the oracle."""
    from oracle import fsnip_cases as fc
    geo = orc.make_geometry(16)
    c0, f, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, 4, seed=5)
    f = f.copy()
    f[1] = 0.0
    for group, ngroups in ((0, 1), (1, 2)):
        a = fc.rhoofr_kpt(geo, c0, f, 0.37, hgkp, hgkm, 2.2, 1.1, group, ngroups)
        b = orc.rhoofr_kpt(geo, c0, f, 0.37, hgkp, hgkm, 2.2, 1.1, group, ngroups)
        assert relmax(b["rhoe"], a["rhoe"]) < 1e-13
        assert abs(a["ekin"] - b["ekin"]) < 1e-12 and abs(a["rsum_g"] - b["rsum_g"]) < 1e-12
        c2a = fc.vpsi_kpt(geo, c0, 0.5 * c0, f, hgkp, hgkm, v, 1.1, group, ngroups)
        assert relmax(orc.vpsi_kpt(geo, c0, 0.5 * c0, f, hgkp, hgkm, v, 1.1, group, ngroups), c2a) < 1e-13


@needs_ref
@pytest.mark.parametrize("geq0", [True, False])
def test_ppener_matches_reference_statements(geq0):
    """ppener_utils.mod.F90:58-104 against the oracle's ppener (both G = 0 branches)."""
    import dataclasses
    from oracle import fsnip_cases as fc
    geo = orc.make_density_geometry((16, 16, 16))
    rng = np.random.default_rng(3)
    n = geo.ngw
    rhog, eivps, eirop = (rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(3))
    scg = rng.random(n) + 0.1
    g = dataclasses.replace(geo, geq0=geq0) if dataclasses.is_dataclass(geo) else geo
    want = orc.ppener(g, rhog, scg, eivps, eirop)
    got = fc.ppener(g, rhog, scg, eivps, eirop, geq0)
    for w, h in zip(want[:5], got[:5]):
        assert abs(w - h) < 1e-11 * max(1.0, abs(w))
    assert relmax(got[5], want[5]) < 1e-14


@needs_ref
@pytest.mark.parametrize("nsup", [None, 2])
def test_tau_variants_match_reference_statements(nsup):
    """tauofr (tauofr_utils.mod.F90:82-102 with dpsisc :121-135 and tauadd :148-173) and vtaupsi
    (vtaupsi_utils.mod.F90:63-89 with taupot :103-127 and ftauadd :142-163): the reference's state loops executed
    statement by statement, with and without LSD, against the oracle."""
    from oracle import fsnip_cases as fc
    geo = orc.make_geometry(16)
    c0, f, _ = orc.synthetic_inputs(geo, 5, seed=2, f_pattern="mixed")
    gk = orc.gk_cartesian(geo)
    for group, ngroups in ((0, 1), (1, 2)):
        assert relmax(orc.tauofr(geo, c0, f, gk, 2.0, 1.2, nsup, group, ngroups),
                      fc.tauofr(geo, c0, f, gk, 2.0, 1.2, nsup, group, ngroups)) < 1e-13
        vt = np.random.default_rng(1).random((1 if nsup is None else 2, geo.nnr1))
        assert relmax(orc.vtaupsi(geo, c0, 0.3 * c0, f, gk, vt, 1.2, nsup, group, ngroups),
                      fc.vtaupsi(geo, c0, 0.3 * c0, f, gk, vt, 1.2, nsup, group, ngroups)) < 1e-13


@needs_ref
def test_exact_exchange_pair_terms_match_reference_statements():
    """hfxab (hfx_utils.mod.F90:1050-1106, both packings iran = 1, 2) and hfxaa (:1216-1256) against the oracle's
    _hfx_pair / _hfx_diag."""
    from oracle import fsnip_cases as fc
    nr = (16, 16, 16)
    geo_w, geo_d = orc.make_geometry(nr), orc.make_density_geometry(nr)
    c0, _, _ = orc.synthetic_inputs(geo_w, 3, seed=2)
    scgx = orc.hfx_coulomb_kernel(geo_d, 1.1)
    pa = orc.invfftn_sparse(geo_w, orc.set_psi_2_states_g(geo_w, c0[0], c0[1]))
    pb = orc.invfftn_sparse(geo_w, orc.set_psi_2_states_g(geo_w, c0[2], c0[1]))
    for iran in (1, 2):
        e1, a1, b1 = fc.hfxab(geo_w, geo_d, pa, pb, iran, 0.37, scgx, 2.2)
        e2, a2, b2 = orc._hfx_pair(geo_w, geo_d, pa, pb, iran, 0.37, scgx, 2.2)
        assert abs(e1 - e2) < 1e-13 and relmax(a2, a1) < 1e-13 and relmax(b2, b1) < 1e-13
    e1, a1 = fc.hfxaa(geo_w, geo_d, pa, 0.37, scgx, 2.2)
    e2, a2 = orc._hfx_diag(geo_w, geo_d, pa, 0.37, scgx, 2.2)
    assert abs(e1 - e2) < 1e-13 and relmax(a2, a1) < 1e-13


@needs_ref
@pytest.mark.parametrize("nsup", [2, 3])
def test_lsd_variants_match_reference_statements(nsup):
    """rhoofr with the spin-resolved accumulation (rhoofr_utils.mod.F90:369-385, build_density_real / _imag) and vpsi
    with the spin-resolved potential (vpsi_utils.mod.F90:450-482) incl. the pair that straddles the spin boundary."""
    from oracle import fsnip_cases as fc
    geo = orc.make_geometry(16)
    c0, f, _ = orc.synthetic_inputs(geo, 5, seed=2, f_pattern="mixed")
    v2 = np.random.default_rng(3).random((2, geo.nnr1))
    part = orc.rhoofr_lsd(geo, c0, f, 2.0, 1.2, nsup, 0, 2)["rhoe"]          # ngroups = 2: the partial channel densities
    assert relmax(part, fc.rhoofr_lsd(geo, c0, f, 2.0, 1.2, nsup, 0, 2)) < 1e-13
    assert relmax(orc.vpsi_lsd(geo, c0, 0.3 * c0, f, v2, 1.2, nsup), fc.vpsi_lsd(geo, c0, 0.3 * c0, f, v2, 1.2, nsup)) < 1e-13


# ---------------------------------------------------------------------------------------------------------------
# The golden vectors the simulator and GPU tests use (tests/golden/, made from the oracle by tools/make_golden.py)
# reproduced by the reference's statements: links those fixtures to reference code, not only to the restatement.
# ---------------------------------------------------------------------------------------------------------------
@needs_ref
def test_committed_golden_vectors_are_reproduced_by_reference_statements():
    from helpers import (golden_cases, golden_kpt_cases, golden_lsd_cases, golden_tau_cases, load_golden,
                         load_golden_kpt, load_golden_lsd, load_golden_tau)
    from oracle import fsnip_cases as fc
    for path in golden_cases():
        d = load_golden(path)
        geo = _geo(d)
        assert relmax(fc.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["group"], d["ngroups"]), d["rhoe"]) < 1e-13
        assert relmax(fc.vpsi(geo, d["c0"], d["c2_in"], d["f"], d["vpot"], d["tpiba2"], d["group"], d["ngroups"]),
                      d["c2_out"]) < 1e-13
        if d["ngroups"] == 1:
            ekin, rsum = fc.kin_energy(geo, d["c0"], d["f"], d["tpiba2"])
            assert abs(ekin - d["ekin"]) < 1e-12 * max(1.0, abs(ekin)) and abs(rsum - d["rsum_g"]) < 1e-12 * max(1.0, rsum)
    for path in golden_lsd_cases():
        d = load_golden_lsd(path)
        geo = _geo(d)
        part = fc.rhoofr_lsd(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["nsup"])
        part[0] += part[1]                                               # alpha+beta / beta (:543-559, pinned separately)
        assert relmax(part, d["rhoe"]) < 1e-13
        assert relmax(fc.vpsi_lsd(geo, d["c0"], d["c2_in"], d["f"], d["vpot"], d["tpiba2"], d["nsup"]), d["c2_out"]) < 1e-13
    for path in golden_kpt_cases():
        d = load_golden_kpt(path)
        geo = _geo(d)
        r = fc.rhoofr_kpt(geo, d["c0"], d["f"], d["wk"], d["hgkp"], d["hgkm"], d["omega"], d["tpiba2"])
        assert relmax(r["rhoe"], d["rhoe"]) < 1e-13 and abs(r["ekin"] - d["ekin"]) < 1e-12 and abs(r["rsum_g"] - d["rsum_g"]) < 1e-12
        assert relmax(fc.vpsi_kpt(geo, d["c0"], d["c2_in"], d["f"], d["hgkp"], d["hgkm"], d["vpot"], d["tpiba2"]),
                      d["c2_out"]) < 1e-13
    for path in golden_tau_cases():
        d = load_golden_tau(path)
        geo = _geo(d)
        nsup = None if d["nsup"] < 0 else d["nsup"]
        assert relmax(fc.tauofr(geo, d["c0"], d["f"], d["gk"], d["omega"], d["tpiba2"], nsup), d["tau"]) < 1e-13
        assert relmax(fc.vtaupsi(geo, d["c0"], d["c2_in"], d["f"], d["gk"], d["vtau"], d["tpiba2"], nsup), d["c2_out"]) < 1e-13


@needs_ref
def test_vofrho_golden_vectors_are_reproduced_by_reference_ppener():
    from helpers import ener_vector, golden_vofrho_cases, load_golden_vofrho
    from oracle import fsnip_cases as fc
    for path in golden_vofrho_cases():
        d = load_golden_vofrho(path)
        eh, ei, ee, eps, vploc, vtemp = fc.ppener(None, d["rhog"], d["scg"], d["eivps"], d["eirop"], geq0=True)
        assert relmax(vtemp, d["vtemp"]) < 1e-13
        got = ener_vector(dict(eh=complex(eh), ei=complex(ei), ee=complex(ee), eps=complex(eps), vploc=float(vploc)))
        assert np.abs(got - d["ener"]).max() < 1e-11 * max(1.0, np.abs(d["ener"]).max())
