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
