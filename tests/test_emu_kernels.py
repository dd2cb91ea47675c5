"""The CUDA kernel sources, compiled for the CPU functional simulator (tests/emu), against the
oracle.  This is how kernel index arithmetic and the host-side plan/pipeline logic are checked
where no GPU exists; the same comparisons run on the real device in test_gpu_parity.py."""
import numpy as np
import pytest

from cpmd_b200 import lib, synthetic
from cpmd_b200.api import CpbError, CpmdContext, Plan, StopGM
from helpers import ETOL, RTOL, golden_cases, load_golden, relmax
from oracle import cpmd_oracle as orc


def _plan(d, cdll, **kw):
    return Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], _cdll=cdll, **kw)


@pytest.mark.parametrize("n,nstate,mb,fp", [(16, 4, 16, "all2"), (20, 5, 2, "mixed"), (24, 7, 2, "mixed"),
                                            (30, 6, 1, "mixed"), (36, 3, 16, "all2"), (40, 2, 16, "all2"),
                                            (48, 4, 3, "all2"), (60, 2, 16, "all2"), (32, 5, 3, "mixed"),
                                            (64, 3, 2, "mixed")])
def test_kernels_match_oracle(emu_cdll, n, nstate, mb, fp):
    d = synthetic.make_inputs(n, nstate, f_pattern=fp)
    geo = orc.make_geometry(n)
    p = _plan(d, emu_cdll, max_batch=mb)
    nz, iz = p.maps()
    assert np.array_equal(nz, geo.nzhs) and np.array_equal(iz, geo.indzs)
    info = p.info
    assert info["nrays"] == geo.nrays and info["zband"] == geo.kr3max - geo.kr3min + 1
    rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"])
    ref = orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)
    assert relmax(rho, ref["rhoe"]) < RTOL
    assert abs(ekin - ref["ekin"]) < ETOL and abs(rg - ref["rsum_g"]) < ETOL and abs(rr - ref["rsum_r"]) < ETOL
    c2 = 0.5 * d["c0"]
    c2_ref = orc.vpsi(geo, d["c0"], c2, d["f"], d["vpot"], 1.0)
    p.vpsi(d["c0"], c2, d["f"], d["vpot"])
    assert relmax(c2, c2_ref) < RTOL
    # pads of rho are exactly zero
    r3 = rho.reshape(p.kr[2], p.kr[1], p.kr[0])
    assert not r3[n:].any() and not r3[:, n:].any() and not r3[:, :, n:].any()


@pytest.mark.parametrize("n,radix", [(16, (4, 4)), (32, (8, 4)), (48, (12, 4)), (64, (8, 8))])
def test_warp_z_kernels_selected_and_match_block_kernels(emu_cdll, monkeypatch, n, radix):
    """The warp-autonomous z kernels (kernels_zw.h, selected with CPB_ZW=1 where the length has a CPB_ZW
    factorisation and the band fits) against the block kernels and the oracle, incl. odd state counts and
    several batches."""
    d = synthetic.make_inputs(n, 5, f_pattern="mixed")
    monkeypatch.setenv("CPB_ZW", "1")
    p = _plan(d, emu_cdll, max_batch=2)
    assert p.info["z_warp_kernels"] and p.info["z_warp_radix"] == radix
    monkeypatch.setenv("CPB_ZW", "0")
    q = _plan(d, emu_cdll, max_batch=2)
    monkeypatch.delenv("CPB_ZW")
    assert not q.info["z_warp_kernels"] and q.info["z_warp_radix"] == (0, 0)
    geo = orc.make_geometry(n)
    rho_w, ekin_w, *_ = p.rhoofr(d["c0"], d["f"])
    assert relmax(rho_w, orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)["rhoe"]) < RTOL
    rho_b, ekin_b, *_ = q.rhoofr(d["c0"], d["f"])
    assert relmax(rho_w, rho_b) < 1e-13 and abs(ekin_w - ekin_b) < ETOL
    c2w, c2b = 0.5 * d["c0"], 0.5 * d["c0"]
    p.vpsi(d["c0"], c2w, d["f"], d["vpot"])
    q.vpsi(d["c0"], c2b, d["f"], d["vpot"])
    assert relmax(c2w, c2b) < 1e-13
    assert relmax(c2w, orc.vpsi(geo, d["c0"], 0.5 * d["c0"], d["f"], d["vpot"], 1.0)) < RTOL


def test_default_batch_is_sized_from_the_mesh(emu_cdll):
    """max_batch_pairs <= 0: a multiple of 8 pairs in [8, 64] (about 3 GB of work space); small meshes get 64."""
    d = synthetic.make_inputs(16, 3)
    assert _plan(d, emu_cdll, max_batch=0).info["max_batch"] == 64
    assert _plan(d, emu_cdll, max_batch=5).info["max_batch"] == 5
    p = _plan(d, emu_cdll, max_batch=0)
    rho, *_ = p.rhoofr(d["c0"], d["f"])
    assert relmax(rho, orc.rhoofr(orc.make_geometry(16), d["c0"], d["f"], 1.0, 1.0)["rhoe"]) < RTOL


def test_warp_z_kernels_not_used_without_factorisation(emu_cdll, monkeypatch):
    monkeypatch.setenv("CPB_ZW", "1")
    d = synthetic.make_inputs(20, 2)
    assert not _plan(d, emu_cdll).info["z_warp_kernels"]


def _stretched(nx, nyz):
    """(nx, nyz, nyz) mesh whose cutoff ellipsoid fills the middle half of the long x axis: b1 shortened so that
    |G|^2 < (nyz/4)^2 reaches x index nx/4 (exercises every band position of the x kernels at small cost)."""
    nr = (nx, nyz, nyz)
    b = np.diag([nyz / float(nx), 1.0, 1.0])
    return nr, orc.make_geometry(nr, gcutw=(nyz / 4.0) ** 2, b=b)


@pytest.mark.parametrize("nr_b,xw,ra", [((64, 64, 64), 1, 8), ("s128", 1, 16), ("s192", 1, 24), ("s192", 3, 24),
                                        ((64, 64, 64), 3, 8)])
def test_warp_x_kernels_match_oracle_and_block_kernels(emu_cdll, monkeypatch, nr_b, xw, ra):
    """k_xw_inv / k_xw_fwd (kernels_xw.h, CPB_XW bit 0 / bit 1) against the oracle and the block mirror kernels:
    rho, kin_energy / dotp sums folded into the gather, C2 with the += semantics, odd state count, several
    batches and pair groups."""
    if isinstance(nr_b, str):
        nr, geo = _stretched(int(nr_b[1:]), 16)
    else:
        nr, geo = nr_b, orc.make_geometry(nr_b)
    c0, f, v = orc.synthetic_inputs(geo, 5, f_pattern="mixed")
    monkeypatch.setenv("CPB_XW", str(xw))
    p = Plan(nr, geo.inyh, geo.hg, 1.0, 1.0, max_batch=2, _cdll=emu_cdll)
    monkeypatch.setenv("CPB_XW", "0")
    q = Plan(nr, geo.inyh, geo.hg, 1.0, 1.0, max_batch=2, _cdll=emu_cdll)
    monkeypatch.delenv("CPB_XW")
    assert p.info["x_warp_kernels"] == xw and p.info["x_warp_radix"] == ra and q.info["x_warp_kernels"] == 0
    rho_w, ekin_w, rg_w, rr_w = p.rhoofr(c0, f)
    rho_b, ekin_b, rg_b, rr_b = q.rhoofr(c0, f)
    ref = orc.rhoofr(geo, c0, f, 1.0, 1.0)
    assert relmax(rho_w, ref["rhoe"]) < RTOL and relmax(rho_w, rho_b) < 1e-13
    assert abs(ekin_w - ref["ekin"]) < ETOL * max(1.0, abs(ref["ekin"])) and abs(rg_w - ref["rsum_g"]) < ETOL
    c2w, c2b = 0.5 * c0, 0.5 * c0
    p.vpsi(c0, c2w, f, v)
    q.vpsi(c0, c2b, f, v)
    assert relmax(c2w, orc.vpsi(geo, c0, 0.5 * c0, f, v, 1.0)) < RTOL and relmax(c2w, c2b) < 1e-13


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_kernels_match_golden(emu_cdll, path):
    d = load_golden(path)
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=2, _cdll=emu_cdll)
    rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"], ngroups=d["ngroups"], my_group=d["group"])
    assert relmax(rho, d["rhoe"]) < RTOL
    if d["ngroups"] == 1:
        assert abs(ekin - d["ekin"]) < ETOL and abs(rg - d["rsum_g"]) < ETOL
    assert abs(rr - d["rsum_r"]) < ETOL
    c2 = d["c2_in"].copy()
    p.vpsi(d["c0"], c2, d["f"], d["vpot"], ngroups=d["ngroups"], my_group=d["group"])
    assert relmax(c2, d["c2_out"]) < RTOL


def test_overwrite_tksham_and_leading_dimension(emu_cdll):
    n, ns = 16, 5
    d = synthetic.make_inputs(n, ns, f_pattern="mixed")
    geo = orc.make_geometry(n)
    p = _plan(d, emu_cdll, max_batch=2)
    ld = geo.ngw + 7                       # ld > ngw as in c0(nkpt%ngwk, nstate)
    c0 = np.zeros((ns, ld), complex)
    c0[:, :geo.ngw] = d["c0"]
    c2 = np.full((ns, ld), 3.0 + 1j)
    p.vpsi(c0, c2, d["f"], d["vpot"], flags=lib.CPB_VPSI_OVERWRITE | lib.CPB_VPSI_TKSHAM)
    ref = orc.vpsi(geo, d["c0"], np.zeros_like(d["c0"]), d["f"], d["vpot"], 1.0, tksham=True)
    assert relmax(c2[:, :geo.ngw], ref) < RTOL
    assert np.all(c2[:, geo.ngw:] == 3.0 + 1j)          # padding rows untouched
    rho, *_ = p.rhoofr(c0, d["f"])
    assert relmax(rho, orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)["rhoe"]) < RTOL


def test_groups_sum_to_full_and_c0_cache(emu_cdll):
    n, ns = 20, 7
    d = synthetic.make_inputs(n, ns, f_pattern="mixed")
    geo = orc.make_geometry(n)
    p = _plan(d, emu_cdll, max_batch=2)
    full = orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)
    c2_full = orc.vpsi(geo, d["c0"], np.zeros_like(d["c0"]), d["f"], d["vpot"], 1.0)
    acc = np.zeros(p.nnr1)
    sums = np.zeros(3)
    c2 = np.zeros_like(d["c0"])
    for grp in range(3):
        rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"], ngroups=3, my_group=grp, flags=lib.CPB_C0_KEEP)
        acc += rho
        sums += (ekin, rg, rr)
        p.vpsi(d["c0"], c2, d["f"], d["vpot"], ngroups=3, my_group=grp, flags=lib.CPB_C0_REUSE)
    assert relmax(acc, full["rhoe"]) < RTOL
    assert np.abs(sums - (full["ekin"], full["rsum_g"], full["rsum_r"])).max() < ETOL
    assert relmax(c2, c2_full) < RTOL
    # a stale cache key must not be reused for a different array
    other = d["c0"] * 2.0
    rho2, *_ = p.rhoofr(other, d["f"], flags=lib.CPB_C0_REUSE)
    assert relmax(rho2, 4.0 * full["rhoe"]) < RTOL


def test_anisotropic_mesh(emu_cdll):
    nr = (16, 20, 24)
    geo = orc.make_geometry(nr)
    c0, f, v = orc.synthetic_inputs(geo, 3)
    p = Plan(nr, geo.inyh, geo.hg, 1.0, 1.0, max_batch=2, _cdll=emu_cdll)
    rho, *_ = p.rhoofr(c0, f)
    assert relmax(rho, orc.rhoofr(geo, c0, f, 1.0, 1.0)["rhoe"]) < RTOL
    c2 = np.zeros_like(c0)
    p.vpsi(c0, c2, f, v)
    assert relmax(c2, orc.vpsi(geo, c0, np.zeros_like(c0), f, v, 1.0)) < RTOL


def test_shuffled_g_order(emu_cdll):
    """Neither routine depends on the order inside a |G|^2 shell: permuting the
    plane waves (G=0 kept first) permutes c2 and leaves rho unchanged."""
    n, ns = 16, 4
    d = synthetic.make_inputs(n, ns)
    rng = np.random.default_rng(7)
    perm = np.concatenate([[0], 1 + rng.permutation(d["hg"].shape[0] - 1)])
    p0 = _plan(d, emu_cdll)
    p1 = Plan(d["nr"], d["inyh"][:, perm], d["hg"][perm], 1.0, 1.0, _cdll=emu_cdll)
    c0p = np.ascontiguousarray(d["c0"][:, perm])
    r0, *_ = p0.rhoofr(d["c0"], d["f"])
    r1, *_ = p1.rhoofr(c0p, d["f"])
    assert relmax(r1, r0) < RTOL
    a = np.zeros_like(d["c0"])
    b = np.zeros_like(d["c0"])
    p0.vpsi(d["c0"], a, d["f"], d["vpot"])
    p1.vpsi(c0p, b, d["f"], d["vpot"])
    assert relmax(b, a[:, perm]) < RTOL


def test_empty_and_single_state(emu_cdll):
    d = synthetic.make_inputs(16, 1)
    geo = orc.make_geometry(16)
    p = _plan(d, emu_cdll)
    rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"])
    assert relmax(rho, orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)["rhoe"]) < RTOL
    # a group that owns no state (nstate < ngroups): rho = 0, scalars = 0
    rho, ekin, rg, rr = p.rhoofr(d["c0"], d["f"], ngroups=2, my_group=1)
    assert not rho.any() and ekin == 0.0 and rg == 0.0 and rr == 0.0
    c2 = np.ones_like(d["c0"])
    p.vpsi(d["c0"], c2, d["f"], d["vpot"], ngroups=2, my_group=1)
    assert np.all(c2 == 1.0)
    # all occupations zero: every pair is skipped in rhoofr (rhoofr_utils.mod.F90:312-316)
    rho, ekin, rg, rr = p.rhoofr(d["c0"], np.zeros(1))
    assert not rho.any() and ekin == 0.0


def test_error_paths(emu_cdll):
    d = synthetic.make_inputs(16, 2)
    with pytest.raises(CpbError) as e:
        Plan((22, 22, 22), d["inyh"], d["hg"], _cdll=emu_cdll)          # no kernel for 22
    assert e.value.code == lib.CPB_ERR_UNSUPPORTED
    bad = d["inyh"].copy()
    bad[0, 5] = 1                                                        # mirror falls outside the mesh
    with pytest.raises(CpbError) as e:
        Plan(d["nr"], bad, d["hg"], _cdll=emu_cdll)
    assert e.value.code == lib.CPB_ERR_INVALID
    dup = d["inyh"].copy()
    dup[:, 9] = dup[:, 8]
    with pytest.raises(CpbError):
        Plan(d["nr"], dup, d["hg"], _cdll=emu_cdll)
    both = d["inyh"].copy()
    both[:, 9] = 2 * 9 - both[:, 8]                                      # -G of entry 8
    with pytest.raises(CpbError):
        Plan(d["nr"], both, d["hg"], _cdll=emu_cdll)
    p = _plan(d, emu_cdll)
    with pytest.raises(ValueError):
        p.rhoofr(d["c0"][:, :50], d["f"])                                # ld < ngw
    with pytest.raises(CpbError):
        p.rhoofr(d["c0"], d["f"], ngroups=2, my_group=2)


def test_context_mirrors_reference_signatures(emu_cdll):
    n, ns = 16, 4
    d = synthetic.make_inputs(n, ns)
    geo = orc.make_geometry(n)
    ctx = CpmdContext(nr=d["nr"], inyh=d["inyh"], hg=d["hg"], f=d["f"], _cdll=emu_cdll)
    rhoe = np.zeros((ctx.nnr1, 1))
    psi = np.zeros(1, complex)
    ctx.rhoofr(d["c0"], rhoe, psi, ns)
    ref = orc.rhoofr(geo, d["c0"], d["f"], 1.0, 1.0)
    assert relmax(rhoe[:, 0], ref["rhoe"]) < RTOL and abs(ctx.ekin - ref["ekin"]) < ETOL
    assert abs(ctx.csumg - ctx.csumr) < 1e-10
    c2 = np.zeros_like(d["c0"])
    ctx.vpsi(d["c0"], c2, d["f"], d["vpot"].reshape(-1, 1), psi, ns, 1, 1, False)
    assert relmax(c2, orc.vpsi(geo, d["c0"], np.zeros_like(c2), d["f"], d["vpot"], 1.0)) < RTOL
    # unsupported variants stop like the reference
    with pytest.raises(StopGM):
        ctx.vpsi(d["c0"], c2, d["f"], d["vpot"], psi, ns, ikind=2)
    ctx.tlse = True
    with pytest.raises(StopGM):
        ctx.rhoofr(d["c0"], rhoe, psi, ns)
    ctx.tlse = False
    # charge check (rhoofr_utils.mod.F90:625-635): a non-normalisable input still passes the
    # identity, so provoke it by lying about omega-independent sums via delta
    ctx.delta = -1.0
    with pytest.raises(StopGM):
        ctx.rhoofr(d["c0"], rhoe, psi, ns)


def test_two_work_spaces(emu_cdll, monkeypatch):
    """Two work spaces / streams (the default): consecutive batches alternate between them; results are
    bit-identical to the single-stream run, CPB_STREAMS=1 (rho is still accumulated batch after batch in order)."""
    d = synthetic.make_inputs(20, 9, f_pattern="mixed")
    monkeypatch.setenv("CPB_STREAMS", "1")
    p1 = _plan(d, emu_cdll, max_batch=2)
    assert p1.info["streams"] == 1
    monkeypatch.delenv("CPB_STREAMS")
    p2 = _plan(d, emu_cdll, max_batch=2)
    assert p2.info["streams"] == 2
    r1, *s1 = p1.rhoofr(d["c0"], d["f"])
    r2, *s2 = p2.rhoofr(d["c0"], d["f"])
    assert np.array_equal(r1, r2) and s1 == s2
    a = 0.25 * d["c0"]
    b = a.copy()
    p1.vpsi(d["c0"], a, d["f"], d["vpot"])
    p2.vpsi(d["c0"], b, d["f"], d["vpot"])
    assert np.array_equal(a, b)
    p2.set_streams(1)
    assert p2.info["streams"] == 1
    with pytest.raises(CpbError):
        p1.set_streams(2)                  # only one work space was allocated


@pytest.mark.parametrize("n,nstate,nsup,mb", [(16, 7, 3, 2), (16, 7, 4, 2), (20, 6, 0, 16), (20, 6, 6, 16),
                                              (24, 9, 5, 1), (16, 2, 1, 16)])
def test_lsd_matches_oracle(emu_cdll, n, nstate, nsup, mb):
    """cntl%tlsd: rhoofr_utils.mod.F90:375-385,543-559 and vpsi_utils.mod.F90:450-482."""
    d = synthetic.make_inputs(n, nstate, f_pattern="mixed")
    geo = orc.make_geometry(n)
    p = _plan(d, emu_cdll, max_batch=mb)
    ref = orc.rhoofr_lsd(geo, d["c0"], d["f"], 1.0, 1.0, nsup)
    rho, ekin, rg, rr, cs, ca = p.rhoofr_lsd(d["c0"], d["f"], nsup)
    scale = np.abs(ref["rhoe"][0]).max()
    assert np.abs(rho - ref["rhoe"]).max() / scale < RTOL
    assert abs(ekin - ref["ekin"]) < ETOL and abs(rg - ref["rsum_g"]) < ETOL and abs(rr - ref["rsum_r"]) < ETOL
    assert abs(cs - ref["csums"]) < ETOL and abs(ca - ref["csumsabs"]) < ETOL
    v2 = np.stack([d["vpot"], 0.5 * d["vpot"][::-1]])
    c2 = 0.5 * d["c0"]
    c2_ref = orc.vpsi_lsd(geo, d["c0"], c2, d["f"], v2, 1.0, nsup)
    p.vpsi_lsd(d["c0"], c2, d["f"], nsup, v2)
    assert relmax(c2, c2_ref) < RTOL


def test_lsd_groups_and_context(emu_cdll):
    n, ns, nsup = 16, 9, 4
    d = synthetic.make_inputs(n, ns, f_pattern="mixed")
    geo = orc.make_geometry(n)
    p = _plan(d, emu_cdll, max_batch=2)
    v2 = np.stack([d["vpot"], 0.25 * d["vpot"]])
    full = orc.rhoofr_lsd(geo, d["c0"], d["f"], 1.0, 1.0, nsup)
    c2_full = orc.vpsi_lsd(geo, d["c0"], np.zeros_like(d["c0"]), d["f"], v2, 1.0, nsup)
    acc = np.zeros((2, p.nnr1))
    c2 = np.zeros_like(d["c0"])
    for grp in range(3):
        rho, *_ = p.rhoofr_lsd(d["c0"], d["f"], nsup, ngroups=3, my_group=grp)
        part = orc.rhoofr_lsd(geo, d["c0"], d["f"], 1.0, 1.0, nsup, group=grp, ngroups=3)["rhoe"]
        assert np.abs(rho - part).max() < RTOL * np.abs(full["rhoe"][0]).max()    # raw partial channels
        acc += rho
        p.vpsi_lsd(d["c0"], c2, d["f"], nsup, v2, ngroups=3, my_group=grp)
    rr, cs, ca = orc.lsd_finish(geo, acc, 1.0)                                   # after cp_grp_redist
    assert np.abs(acc - full["rhoe"]).max() < RTOL * np.abs(full["rhoe"][0]).max()
    assert abs(cs - full["csums"]) < ETOL and abs(ca - full["csumsabs"]) < ETOL and abs(rr - full["rsum_r"]) < ETOL
    assert relmax(c2, c2_full) < RTOL
    # the context with cntl%tlsd mirrors rhoe(nnr1,2) / vpot(nnr1,2)
    ctx = CpmdContext(nr=d["nr"], inyh=d["inyh"], hg=d["hg"], f=d["f"], tlsd=True, nsup=nsup, _cdll=emu_cdll)
    rhoe = np.zeros((2, ctx.nnr1))
    ctx.rhoofr(d["c0"], rhoe, None, ns)
    assert np.abs(rhoe - full["rhoe"]).max() < RTOL * np.abs(full["rhoe"][0]).max()
    assert abs(ctx.csums - full["csums"]) < ETOL and abs(ctx.csumsabs - full["csumsabs"]) < ETOL
    c2b = np.zeros_like(d["c0"])
    ctx.vpsi(d["c0"], c2b, d["f"], v2, None, ns, 1, 2, False)
    assert relmax(c2b, c2_full) < RTOL
    with pytest.raises(CpbError):
        p.rhoofr_lsd(d["c0"], d["f"], ns + 1)


def test_psi_keep_reuse(emu_cdll):
    """CPB_PSI_KEEP / CPB_PSI_REUSE (device-side REAL SPACE WFN KEEP): vpsi starts from the y-pass
    output rhoofr left behind; same kernels on the same data, so the result is bit-identical."""
    d = synthetic.make_inputs(20, 7, f_pattern="mixed")          # has unoccupied states
    p = _plan(d, emu_cdll, max_batch=2)
    rho0, *s0 = p.rhoofr(d["c0"], d["f"])
    a = 0.5 * d["c0"]
    p.vpsi(d["c0"], a, d["f"], d["vpot"])
    rho1, *s1 = p.rhoofr(d["c0"], d["f"], flags=lib.CPB_C0_KEEP | lib.CPB_PSI_KEEP)
    assert np.array_equal(rho0, rho1) and s0 == s1
    b = 0.5 * d["c0"]
    n0 = p.launch_count
    p.vpsi(d["c0"], b, d["f"], d["vpot"], flags=lib.CPB_C0_REUSE | lib.CPB_PSI_REUSE)
    reuse_launches = p.launch_count - n0
    assert np.array_equal(a, b)
    # the cache is consumed: a second REUSE call runs the full pipeline, same answer
    c = 0.5 * d["c0"]
    n0 = p.launch_count
    p.vpsi(d["c0"], c, d["f"], d["vpot"], flags=lib.CPB_C0_REUSE | lib.CPB_PSI_REUSE)
    assert np.array_equal(a, c) and p.launch_count - n0 > reuse_launches
    # a different array never hits the cache
    p.rhoofr(d["c0"], d["f"], flags=lib.CPB_C0_KEEP | lib.CPB_PSI_KEEP)
    other = d["c0"] * (1.0 + 0.5j)
    e = np.zeros_like(other)
    p.vpsi(other, e, d["f"], d["vpot"], flags=lib.CPB_C0_REUSE | lib.CPB_PSI_REUSE)
    geo = orc.make_geometry(20)
    assert relmax(e, orc.vpsi(geo, other, np.zeros_like(other), d["f"], d["vpot"], 1.0)) < RTOL
    # LSD and groups go through the same cache
    v2 = np.stack([d["vpot"], 0.5 * d["vpot"]])
    ref = np.zeros_like(d["c0"])
    got = np.zeros_like(d["c0"])
    for g in range(2):
        p.vpsi_lsd(d["c0"], ref, d["f"], 3, v2, ngroups=2, my_group=g)
        p.rhoofr_lsd(d["c0"], d["f"], 3, ngroups=2, my_group=g, flags=lib.CPB_C0_KEEP | lib.CPB_PSI_KEEP)
        p.vpsi_lsd(d["c0"], got, d["f"], 3, v2, ngroups=2, my_group=g, flags=lib.CPB_C0_REUSE | lib.CPB_PSI_REUSE)
    assert np.array_equal(ref, got)


# ---------------------------------------------------------------------------------------------
# dense transforms on the density cutoff + local part of vofrho
# ---------------------------------------------------------------------------------------------
from helpers import ener_vector, golden_vofrho_cases, load_golden_vofrho, padded_random  # noqa: E402


def _dense_plan(nr, cdll, **kw):
    geo = orc.make_density_geometry(nr)
    return geo, Plan(geo.nr, geo.inyh, geo.hg, 1.0, 1.0, _cdll=cdll, max_batch=2, **kw)


@pytest.mark.parametrize("nr", [16, 20, (16, 20, 24), 30, 36, 48])
def test_vofrho_local_matches_oracle(emu_cdll, nr):
    geo, p = _dense_plan(nr, emu_cdll)
    nz, iz = p.maps()
    assert np.array_equal(nz, geo.nzhs) and np.array_equal(iz, geo.indzs)      # nzh / indz
    if isinstance(nr, int):
        assert p.info["band_pruned"] == (0, 0, 0)    # the sphere fills the box; (16,20,24) prunes z
    rho = padded_random(geo, np.random.default_rng(geo.nr[0]))
    scg, eivps, eirop = orc.synthetic_vofrho_inputs(geo)
    ref = orc.vofrho_local(geo, rho, scg, eivps, eirop)
    rhog = np.empty(geo.ngw, complex)
    vtemp = np.empty(geo.ngw, complex)
    v, e = p.vofrho_local(rho, scg, eivps, eirop, rhog=rhog, vtemp=vtemp)
    assert relmax(rhog, ref["rhog"]) < RTOL and relmax(vtemp, ref["vtemp"]) < RTOL
    assert relmax(v, ref["v"]) < RTOL
    assert np.abs(ener_vector(e) - ener_vector(ref)).max() < ETOL
    n1, n2, n3 = geo.nr
    v3 = v.reshape(geo.kr[2], geo.kr[1], geo.kr[0])
    assert not v3[n3:].any() and not v3[:, n2:].any() and not v3[:, :, n1:].any()
    # in place (v aliases rhoe like the reference's rhoe) and without the optional outputs
    buf = rho.copy()
    v2, e2 = p.vofrho_local(buf, scg, eivps, eirop, v=buf)
    assert np.array_equal(v2, v) and e2 == e


@pytest.mark.parametrize("path", golden_vofrho_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_vofrho_local_matches_golden(emu_cdll, path):
    d = load_golden_vofrho(path)
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=1, _cdll=emu_cdll)
    rhog = np.empty(p.ngw, complex)
    vtemp = np.empty(p.ngw, complex)
    v, e = p.vofrho_local(d["rhoe"], d["scg"], d["eivps"], d["eirop"], rhog=rhog, vtemp=vtemp)
    assert relmax(v, d["v"]) < RTOL and relmax(rhog, d["rhog"]) < RTOL and relmax(vtemp, d["vtemp"]) < RTOL
    assert np.abs(ener_vector(e) - d["ener"]).max() < ETOL


def test_dense_transforms_one_and_two_fields(emu_cdll):
    """cpb_dense_fwfft_dev / cpb_dense_invfft_dev (the simulator's "device" is host memory): one
    field, two fields packed into one transform, leading dimension > nhg, accumulate flag."""
    geo, p = _dense_plan(20, emu_cdll)
    rng = np.random.default_rng(3)
    f2 = np.stack([padded_random(geo, rng), padded_random(geo, rng)])
    ld = geo.ngw + 5
    g2 = np.full((2, ld), 9.0 + 9.0j)
    p.dense_fwfft_dev(f2, g2)
    for i in range(2):
        assert relmax(g2[i, :geo.ngw], orc.rho_to_g(geo, f2[i])) < RTOL
    assert np.all(g2[:, geo.ngw:] == 9.0 + 9.0j)
    g1 = np.empty(ld, complex)
    p.dense_fwfft_dev(f2[1], g1)
    assert relmax(g1[:geo.ngw], orc.rho_to_g(geo, f2[1])) < RTOL
    # inverse: two fields at once == one at a time == oracle
    back2 = np.full((2, geo.nnr1), 5.0)
    p.dense_invfft_dev(g2, back2)
    for i in range(2):
        assert relmax(back2[i], orc.g_to_r(geo, g2[i, :geo.ngw]).real) < RTOL
    back1 = np.full(geo.nnr1, 5.0)
    p.dense_invfft_dev(g2[0], back1)
    assert relmax(back1, back2[0]) < RTOL
    acc = back1.copy()
    p.dense_invfft_dev(g2[0], acc, accumulate=True)
    assert relmax(acc, 2.0 * back1) < RTOL
    # argument checks
    with pytest.raises(ValueError):
        p.dense_fwfft_dev(f2[0], np.empty(geo.ngw - 1, complex))
    with pytest.raises(CpbError):
        p.dense_fwfft_dev(np.empty((3, geo.nnr1)), np.empty((3, ld), complex))


def test_density_to_potential_step(emu_cdll):
    """rhoofr -> vofrho_local -> vpsi chained like one SCF step (rwfopt_utils.mod.F90:329 ->
    vofrho -> forces_driver.mod.F90:224), wavefunction plan + density plan on the same mesh."""
    n, ns = 16, 4
    d = synthetic.make_inputs(n, ns)
    wgeo = orc.make_geometry(n)
    wp = _plan(d, emu_cdll, max_batch=2)
    dgeo, dp = _dense_plan(n, emu_cdll)
    scg, eivps, eirop = orc.synthetic_vofrho_inputs(dgeo)
    rho, *_ = wp.rhoofr(d["c0"], d["f"])
    v, e = dp.vofrho_local(rho, scg, eivps, eirop, v=rho)           # rho becomes V in place
    c2 = np.zeros_like(d["c0"])
    wp.vpsi(d["c0"], c2, d["f"], v)
    rho_ref = orc.rhoofr(wgeo, d["c0"], d["f"], 1.0, 1.0)["rhoe"]
    ref = orc.vofrho_local(dgeo, rho_ref, scg, eivps, eirop)
    c2_ref = orc.vpsi(wgeo, d["c0"], np.zeros_like(d["c0"]), d["f"], ref["v"], 1.0)
    assert relmax(v, ref["v"]) < RTOL and relmax(c2, c2_ref) < RTOL
    assert np.abs(ener_vector(e) - ener_vector(ref)).max() < ETOL


# ---------------------------------------------------------------------------------------------
# k-points: one k-point of rhoofr_c and of vpsi's k-branch
# ---------------------------------------------------------------------------------------------
from helpers import golden_kpt_cases, load_golden_kpt  # noqa: E402


@pytest.mark.parametrize("nr,ns,mb", [(16, 5, 2), (20, 4, 3), ((16, 20, 24), 3, 16), (30, 6, 1), (36, 2, 16)])
def test_kpt_matches_oracle(emu_cdll, nr, ns, mb):
    geo = orc.make_geometry(nr)
    p = Plan(geo.nr, geo.inyh, geo.hg, 0.9, 1.3, max_batch=mb, _cdll=emu_cdll)
    c0, f, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, ns)
    rho = np.full(geo.nnr1, 7.0)                                   # must be zeroed by the first k-point
    ek, rg, rr = p.rhoofr_kpt_dev(c0, f, 0.4, hgkp, hgkm, rho)
    ref = orc.rhoofr_kpt(geo, c0, f, 0.4, hgkp, hgkm, 1.3, 0.9)
    assert relmax(rho, ref["rhoe"]) < RTOL
    assert abs(ek - ref["ekin"]) < ETOL and abs(rg - ref["rsum_g"]) < ETOL and abs(rr - rg) < ETOL
    # second k-point accumulates (rhoofr_c zeroes rhoe once, rhoofr_c_utils.mod.F90:107)
    hg2p, hg2m = hgkm, hgkp                                        # k -> -k
    ek2, rg2, rr2 = p.rhoofr_kpt_dev(c0[::-1].copy(), f, 0.6, hg2p, hg2m, rho, accumulate=True)
    ref2 = orc.rhoofr_kpt(geo, c0[::-1], f, 0.6, hg2p, hg2m, 1.3, 0.9, rhoe=ref["rhoe"].copy())
    assert relmax(rho, ref2["rhoe"]) < RTOL and abs(rr2 - (rg + rg2)) < ETOL
    assert abs(ek2 - ref2["ekin"]) < ETOL
    # vpsi k-branch: += and overwrite, ld > 2 ngw
    ld = 2 * geo.ngw + 3
    c0p = np.zeros((ns, ld), complex)
    c0p[:, :2 * geo.ngw] = c0
    c2 = np.full((ns, ld), 0.5 - 0.25j)
    c2_ref = orc.vpsi_kpt(geo, c0, c2[:, :2 * geo.ngw].copy(), f, hgkp, hgkm, v, 0.9)
    p.vpsi_kpt_dev(c0p, c2, f, hgkp, hgkm, v)
    assert relmax(c2[:, :2 * geo.ngw], c2_ref) < RTOL
    assert np.all(c2[:, 2 * geo.ngw:] == 0.5 - 0.25j)
    p.vpsi_kpt_dev(c0p, c2, f, hgkp, hgkm, v, flags=lib.CPB_VPSI_OVERWRITE)
    assert relmax(c2[:, :2 * geo.ngw], orc.vpsi_kpt(geo, c0, np.zeros_like(c0), f, hgkp, hgkm, v, 0.9)) < RTOL
    # groups add up / partition the states
    acc = np.zeros(geo.nnr1)
    part = np.empty(geo.nnr1)
    c2g = np.zeros_like(c0)
    for g in range(2):
        p.rhoofr_kpt_dev(c0, f, 0.4, hgkp, hgkm, part, ngroups=2, my_group=g)
        acc += part
        p.vpsi_kpt_dev(c0, c2g, f, hgkp, hgkm, v, ngroups=2, my_group=g)
    assert relmax(acc, ref["rhoe"]) < RTOL
    assert relmax(c2g, orc.vpsi_kpt(geo, c0, np.zeros_like(c0), f, hgkp, hgkm, v, 0.9)) < RTOL
    # the Gamma entry points of the same plan are unaffected by the k-point calls
    d = synthetic.make_inputs(geo.nr[0], 2) if isinstance(nr, int) else None
    if d is not None:
        pg = _plan(d, emu_cdll, max_batch=mb)
        r0, *_ = pg.rhoofr(d["c0"], d["f"])
        pg.rhoofr_kpt_dev(c0, f, 0.4, hgkp, hgkm, part)
        r1, *_ = pg.rhoofr(d["c0"], d["f"])
        assert np.array_equal(r0, r1)


@pytest.mark.parametrize("path", golden_kpt_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_kpt_matches_golden(emu_cdll, path):
    d = load_golden_kpt(path)
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=2, _cdll=emu_cdll)
    rho = np.empty(p.nnr1)
    ek, rg, rr = p.rhoofr_kpt_dev(d["c0"], d["f"], d["wk"], d["hgkp"], d["hgkm"], rho)
    assert relmax(rho, d["rhoe"]) < RTOL and abs(ek - d["ekin"]) < ETOL and abs(rg - d["rsum_g"]) < ETOL
    c2 = d["c2_in"].copy()
    p.vpsi_kpt_dev(d["c0"], c2, d["f"], d["hgkp"], d["hgkm"], d["vpot"])
    assert relmax(c2, d["c2_out"]) < RTOL


def test_kpt_argument_checks(emu_cdll):
    geo = orc.make_geometry(16)
    p = Plan(geo.nr, geo.inyh, geo.hg, _cdll=emu_cdll)
    c0, f, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, 2)
    with pytest.raises(ValueError):
        p.rhoofr_kpt_dev(c0[:, :geo.ngw].copy(), f, 1.0, hgkp, hgkm, np.empty(geo.nnr1))   # Gamma-sized c0
    with pytest.raises(CpbError):
        p.vpsi_kpt_dev(c0, np.zeros_like(c0), f, None, hgkm, v)


# ---------------------------------------------------------------------------------------------
# meta-GGA tauofr / vtaupsi
# ---------------------------------------------------------------------------------------------
from helpers import golden_tau_cases, load_golden_tau  # noqa: E402


@pytest.mark.parametrize("nr,ns,mb,nsup", [(16, 5, 2, None), (20, 4, 3, None), (16, 7, 2, 3), ((16, 20, 24), 6, 16, 4),
                                           (30, 3, 1, 0), (36, 2, 16, None)])
def test_tau_matches_oracle(emu_cdll, nr, ns, mb, nsup):
    geo = orc.make_geometry(nr)
    p = Plan(geo.nr, geo.inyh, geo.hg, 0.9, 1.3, max_batch=mb, _cdll=emu_cdll)
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    gk = orc.gk_cartesian(geo)
    nl = 1 if nsup is None else 2
    cs = -1 if nsup is None else nsup
    tau = np.full((nl, geo.nnr1), 7.0)
    p.tauofr_dev(c0, f, gk, tau, nsup=cs)
    ref = orc.tauofr(geo, c0, f, gk, 1.3, 0.9, nsup)
    assert np.abs(tau - ref).max() / np.abs(ref).max() < RTOL
    vt = np.ascontiguousarray(np.stack([v, 0.5 * v[::-1]])[:nl])
    c2 = 0.3 * c0
    c2_ref = orc.vtaupsi(geo, c0, c2, f, gk, vt, 0.9, nsup)
    p.vtaupsi_dev(c0, c2, f, gk, vt, nsup=cs)
    assert relmax(c2, c2_ref) < RTOL
    # groups: partial tau adds up, c2 blocks partition
    acc = np.zeros((nl, geo.nnr1))
    part = np.empty((nl, geo.nnr1))
    c2g = 0.3 * c0
    for g in range(2):
        p.tauofr_dev(c0, f, gk, part, nsup=cs, ngroups=2, my_group=g)
        acc += part
        p.vtaupsi_dev(c0, c2g, f, gk, vt, nsup=cs, ngroups=2, my_group=g)
    assert np.abs(acc - ref).max() / np.abs(ref).max() < RTOL and relmax(c2g, c2_ref) < RTOL


@pytest.mark.parametrize("path", golden_tau_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_tau_matches_golden(emu_cdll, path):
    d = load_golden_tau(path)
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=2, _cdll=emu_cdll)
    tau = np.empty_like(d["tau"])
    p.tauofr_dev(d["c0"], d["f"], d["gk"], tau, nsup=d["nsup"])
    assert np.abs(tau - d["tau"]).max() / np.abs(d["tau"]).max() < RTOL
    c2 = d["c2_in"].copy()
    p.vtaupsi_dev(d["c0"], c2, d["f"], d["gk"], np.ascontiguousarray(d["vtau"]), nsup=d["nsup"])
    assert relmax(c2, d["c2_out"]) < RTOL


def test_empty_and_degenerate_inputs(emu_cdll):
    """Ragged / empty inputs the reference loops handle implicitly: no states, one state, a group with
    no states (more groups than states), all occupations zero."""
    geo = orc.make_geometry(16)
    p = Plan(geo.nr, geo.inyh, geo.hg, _cdll=emu_cdll, max_batch=2)
    c0, f, v = orc.synthetic_inputs(geo, 1)
    rho, ek, rg, rr = p.rhoofr(c0, f)
    ref = orc.rhoofr(geo, c0, f, 1.0, 1.0)
    assert relmax(rho, ref["rhoe"]) < RTOL and abs(ek - ref["ekin"]) < ETOL and abs(rg - rr) < ETOL
    c2 = np.zeros_like(c0)
    p.vpsi(c0, c2, f, v)
    assert relmax(c2, orc.vpsi(geo, c0, np.zeros_like(c0), f, v, 1.0)) < RTOL
    c00 = np.zeros((0, geo.ngw), complex)
    rho, ek, rg, rr = p.rhoofr(c00, np.zeros(0))
    assert not rho.any() and (ek, rg, rr) == (0.0, 0.0, 0.0)
    p.vpsi(c00, np.zeros_like(c00), np.zeros(0), v)
    rho, ek, rg, rr = p.rhoofr(c0, f, ngroups=3, my_group=2)          # this group owns no state
    assert not rho.any() and (ek, rg, rr) == (0.0, 0.0, 0.0)
    rho, ek, rg, rr = p.rhoofr(c0, np.zeros(1))                        # nothing occupied: rho = 0
    assert not rho.any() and (ek, rg, rr) == (0.0, 0.0, 0.0)
    c2 = np.zeros_like(c0)
    p.vpsi(c0, c2, np.zeros(1), v)                                     # vpsi still acts (fi = 1)
    assert relmax(c2, orc.vpsi(geo, c0, np.zeros_like(c0), np.zeros(1), v, 1.0)) < RTOL


def test_host_pointer_forms_of_kpt_and_tau(emu_cdll):
    """cpb_rhoofr_kpt / cpb_vpsi_kpt / cpb_tauofr / cpb_vtaupsi (host arrays, what the Fortran shim
    binds): same results as the oracle, groups, ld > rows, LSD, and the Gamma host path afterwards."""
    geo = orc.make_geometry(20)
    p = Plan(geo.nr, geo.inyh, geo.hg, 0.9, 1.3, max_batch=2, _cdll=emu_cdll)
    ns = 5
    c0k, fk, hgkp, hgkm, v = orc.synthetic_kpt_inputs(geo, ns)
    ld = 2 * geo.ngw + 3
    c0p = np.zeros((ns, ld), complex)
    c0p[:, :2 * geo.ngw] = c0k
    rho = np.full(geo.nnr1, 9.0)
    acc_s = np.zeros(2)
    for g in range(2):                       # two groups accumulate into the same host array
        _, ek, rg, rr = p.rhoofr_kpt(c0p, fk, 0.4, hgkp, hgkm, rho, ngroups=2, my_group=g, accumulate=(g > 0))
        acc_s += (ek, rg)
    ref = orc.rhoofr_kpt(geo, c0k, fk, 0.4, hgkp, hgkm, 1.3, 0.9)
    assert relmax(rho, ref["rhoe"]) < RTOL and abs(acc_s[0] - ref["ekin"]) < ETOL and abs(acc_s[1] - ref["rsum_g"]) < ETOL
    c2 = np.full((ns, ld), 0.5 - 0.25j)
    c2_ref = orc.vpsi_kpt(geo, c0k, c2[:, :2 * geo.ngw].copy(), fk, hgkp, hgkm, v, 0.9)
    for g in range(2):
        p.vpsi_kpt(c0p, c2, fk, hgkp, hgkm, v, ngroups=2, my_group=g)
    assert relmax(c2[:, :2 * geo.ngw], c2_ref) < RTOL and np.all(c2[:, 2 * geo.ngw:] == 0.5 - 0.25j)
    # meta-GGA with LSD through the host forms
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    gk = orc.gk_cartesian(geo)
    for nsup in (None, 2):
        cs = -1 if nsup is None else nsup
        nl = 1 if nsup is None else 2
        ref_t = orc.tauofr(geo, c0, f, gk, 1.3, 0.9, nsup)
        tau = p.tauofr(c0, f, gk, nsup=cs)
        assert np.abs(tau - ref_t).max() / np.abs(ref_t).max() < RTOL
        acc = np.zeros_like(ref_t)
        for g in range(3):
            acc += p.tauofr(c0, f, gk, nsup=cs, ngroups=3, my_group=g)
        assert np.abs(acc - ref_t).max() / np.abs(ref_t).max() < RTOL
        vt = np.ascontiguousarray(np.stack([v, 0.5 * v[::-1]])[:nl])
        c2 = 0.3 * c0
        c2_ref = orc.vtaupsi(geo, c0, c2, f, gk, vt, 0.9, nsup)
        for g in range(3):
            p.vtaupsi(c0, c2, f, gk, vt, nsup=cs, ngroups=3, my_group=g)
        assert relmax(c2, c2_ref) < RTOL
    # the Gamma host path (shares the staging buffers) still works afterwards
    rho_g, *_ = p.rhoofr(c0, f)
    assert relmax(rho_g, orc.rhoofr(geo, c0, f, 1.3, 0.9)["rhoe"]) < RTOL


@pytest.mark.parametrize("n,ns", [(16, 3), (20, 4), (24, 5)])
def test_low_dual_cutoff_runs_unpruned_kernels(emu_cdll, n, ns):
    """A wavefunction cutoff with dual < 4 (sphere radius 0.45 n): the coefficient band no longer sits in
    the middle half of the axes, so the plan must select the unpruned (HALF = false) variants of every
    kernel - including k_z_rho / k_z_vpsi, which the dense-plan tests do not reach."""
    geo = orc.make_geometry(n, gcutw=(0.45 * n) ** 2)
    p = Plan(geo.nr, geo.inyh, geo.hg, 0.9, 1.3, max_batch=2, _cdll=emu_cdll)
    assert p.info["band_pruned"] == (0, 0, 0)
    nz, iz = p.maps()
    assert np.array_equal(nz, geo.nzhs) and np.array_equal(iz, geo.indzs)
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    rho, ek, rg, rr = p.rhoofr(c0, f)
    ref = orc.rhoofr(geo, c0, f, 1.3, 0.9)
    assert relmax(rho, ref["rhoe"]) < RTOL and abs(ek - ref["ekin"]) < ETOL * max(1.0, abs(ref["ekin"]))
    c2 = 0.5 * c0
    c2_ref = orc.vpsi(geo, c0, c2, f, v, 0.9)
    p.vpsi(c0, c2, f, v)
    assert relmax(c2, c2_ref) < RTOL


def test_randomised_meshes_cutoffs_and_orders(emu_cdll):
    """Eight seeded random configurations: anisotropic meshes from the simulator's lengths, cutoff radii
    between 0.18 and 0.49 of the shortest axis (pruned, mixed and unpruned kernels; planes without rays;
    ray counts that are not multiples of the x tile), plane waves in a shuffled order (G=0 first),
    odd state counts, random group splits."""
    rng = np.random.default_rng(20261017)
    lengths = [16, 20, 24, 30, 36, 40]
    for case in range(8):
        nr = tuple(int(v) for v in rng.choice(lengths, 3))
        rad = float(rng.uniform(0.18, 0.49)) * min(nr)
        geo = orc.make_geometry(nr, gcutw=rad * rad)
        ns = int(rng.integers(1, 6))
        c0, f, v = orc.synthetic_inputs(geo, ns, seed=case, f_pattern="mixed")
        perm = np.concatenate([[0], 1 + rng.permutation(geo.ngw - 1)]) if geo.ngw > 1 else np.array([0])
        p = Plan(nr, geo.inyh[:, perm], geo.hg[perm], 0.9, 1.3, max_batch=int(rng.integers(1, 4)), _cdll=emu_cdll)
        c0p = np.ascontiguousarray(c0[:, perm])
        ngroups = int(rng.integers(1, 4))
        acc = np.zeros(geo.nnr1)
        c2 = 0.25 * c0p
        for g in range(ngroups):
            rho, *_ = p.rhoofr(c0p, f, ngroups=ngroups, my_group=g)
            acc += rho
            p.vpsi(c0p, c2, f, v, ngroups=ngroups, my_group=g)
        ref = orc.rhoofr(geo, c0, f, 1.3, 0.9)
        tag = f"case {case}: mesh {nr}, radius {rad:.2f}, ngw {geo.ngw}, {ns} states, {ngroups} groups, {p.info['band_pruned']}"
        assert relmax(acc, ref["rhoe"]) < RTOL, tag
        assert relmax(c2, orc.vpsi(geo, c0, 0.25 * c0, f, v, 0.9)[:, perm]) < RTOL, tag


@pytest.mark.parametrize("nr,scale", [(24, 0.62), ((20, 24, 30), 0.6), (48, 0.62)])
def test_general_cell_reciprocal_vectors(emu_cdll, nr, scale):
    """Non-orthorhombic cell: the cutoff region is an oblique ellipsoid in index space
    and hg is not an integer (loadpa_utils.mod.F90:286-335, rggen_utils.mod.F90:121-129)."""
    if isinstance(nr, int):
        nr = (nr, nr, nr)
    b = np.array([[1.0, 0.0, 0.0], [0.27, 1.06, 0.0], [0.14, -0.21, 0.93]])
    geo = orc.make_geometry(nr, gcutw=scale * (min(nr) / 4.0) ** 2, b=b)
    assert np.abs(geo.hg - np.round(geo.hg)).max() > 1e-3
    tpiba2, omega = 0.83, 41.7
    c0, f, v = orc.synthetic_inputs(geo, 5, f_pattern="mixed")
    p = Plan(nr, geo.inyh, geo.hg, tpiba2, omega, max_batch=2, _cdll=emu_cdll)
    nz, iz = p.maps()
    assert np.array_equal(nz, geo.nzhs) and np.array_equal(iz, geo.indzs)
    rho, ekin, rg, rr = p.rhoofr(c0, f)
    ref = orc.rhoofr(geo, c0, f, omega, tpiba2)
    assert relmax(rho, ref["rhoe"]) < RTOL
    assert abs(ekin - ref["ekin"]) < ETOL * max(1.0, abs(ref["ekin"])) and abs(rg - ref["rsum_g"]) < ETOL
    assert abs(rr - ref["rsum_r"]) < ETOL
    c2 = 0.5 * c0
    c2_ref = orc.vpsi(geo, c0, c2, f, v, tpiba2)
    p.vpsi(c0, c2, f, v)
    assert relmax(c2, c2_ref) < RTOL


def test_charge_check_flag(emu_cdll):
    """rhoofr_utils.mod.F90:625-635: |rsum_r - rsum_g| > 1e-6 stops the run (opt-in: CPB_RHO_CHECK_CHARGE)."""
    d = synthetic.make_inputs(16, 4)
    p = _plan(d, emu_cdll, max_batch=2)
    p.rhoofr(d["c0"], d["f"], flags=lib.CPB_RHO_CHECK_CHARGE)
    bad = d["c0"].copy()
    bad[1, 0] = 0.4 + 0.3j                 # G = 0 coefficient not real: not a real function
    with pytest.raises(CpbError) as ei:
        p.rhoofr(bad, d["f"], flags=lib.CPB_RHO_CHECK_CHARGE)
    assert ei.value.code == lib.CPB_ERR_CHARGE
    _, _, rg, rr = p.rhoofr(bad, d["f"])
    assert abs(rg - rr) > 1e-6


# ---------------------------------------------------------------------------------------------
# Hartree-Fock exchange
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nr,ns", [(16, 5), ((16, 20, 24), 4), (24, 3)])
def test_hfx_matches_oracle(emu_cdll, nr, ns):
    gw = orc.make_geometry(nr)
    gd = orc.make_density_geometry(nr)
    tp, om = 0.9, 1.3
    c0, f, _ = orc.synthetic_inputs(gw, ns, f_pattern="mixed")
    scgx = orc.hfx_coulomb_kernel(gd, tp)
    pw = Plan(gw.nr, gw.inyh, gw.hg, tp, om, max_batch=2, _cdll=emu_cdll)
    pd = Plan(gd.nr, gd.inyh, gd.hg, tp, om, max_batch=1, _cdll=emu_cdll)
    c2 = 0.25 * c0
    want, e_ref, v_ref = orc.hfx(gw, gd, c0, c2, f, scgx, om)
    e, v = pw.hfx_dev(pd, c0, c2, f, scgx)
    assert relmax(c2, want) < RTOL
    assert abs(e - e_ref) < ETOL * max(1.0, abs(e_ref)) and abs(v - v_ref) < ETOL * max(1.0, abs(v_ref))
    # the exchange energy is the Euler sum of its own gradient: sum_i dotp(c0_i, dC2_i) = -2 ehfx
    z = np.zeros_like(c0)
    e2, v2 = pw.hfx_dev(pd, c0, z, f, scgx)
    assert abs(v2 + 2.0 * e2) < 1e-10 * abs(e2) and abs(e2 - e) < 1e-12 * abs(e)
    # hybrid prefactor
    z2 = np.zeros_like(c0)
    e3, _ = pw.hfx_dev(pd, c0, z2, f, scgx, pfl=0.25 * 0.2)
    assert abs(e3 - 0.2 * e) < 1e-12 * abs(e) and relmax(z2, 0.2 * z) < 1e-13


def test_hfx_host_form(emu_cdll):
    gw = orc.make_geometry(16)
    gd = orc.make_density_geometry(16)
    c0, f, _ = orc.synthetic_inputs(gw, 4)
    scgx = orc.hfx_coulomb_kernel(gd, 1.0)
    pw = Plan(gw.nr, gw.inyh, gw.hg, max_batch=2, _cdll=emu_cdll)
    pd = Plan(gd.nr, gd.inyh, gd.hg, max_batch=1, _cdll=emu_cdll)
    ld = gw.ngw + 3
    c0p = np.zeros((4, ld), complex)
    c0p[:, :gw.ngw] = c0
    c2 = np.full((4, ld), 1.0 - 2.0j)
    want, e_ref, _ = orc.hfx(gw, gd, c0, c2[:, :gw.ngw], f, scgx, 1.0)
    e, v = pw.hfx(pd, c0p, c2, f, scgx)
    assert relmax(c2[:, :gw.ngw], want) < RTOL and abs(e - e_ref) < ETOL * abs(e_ref)
    assert np.all(c2[:, gw.ngw:] == 1.0 - 2.0j)
