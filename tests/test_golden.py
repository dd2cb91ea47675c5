"""The committed golden vectors (tools/make_golden.py) against both oracle restatements."""
import numpy as np
import pytest

from helpers import ETOL, RTOL, golden_cases, golden_lsd_cases, load_golden, load_golden_lsd, relmax
from oracle import cpmd_oracle as orc
from oracle import staged


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracles_reproduce_golden(path):
    d = load_golden(path)
    geo = orc.fft_maps(d["nr"], d["inyh"], d["hg"])
    assert np.array_equal(geo.nzhs, d["nzhs"]) and np.array_equal(geo.indzs, d["indzs"])
    for impl in (orc, staged):
        r = impl.rhoofr(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["group"], d["ngroups"])
        assert relmax(r["rhoe"], d["rhoe"]) < RTOL
        assert abs(r["ekin"] - d["ekin"]) < ETOL and abs(r["rsum_r"] - d["rsum_r"]) < ETOL
        c2 = impl.vpsi(geo, d["c0"], d["c2_in"], d["f"], d["vpot"], d["tpiba2"], d["group"], d["ngroups"])
        assert relmax(c2, d["c2_out"]) < RTOL


def test_golden_present():
    assert len(golden_cases()) >= 4


@pytest.mark.parametrize("path", golden_lsd_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_and_simulator_reproduce_lsd_golden(emu_cdll, path):
    from cpmd_b200.api import Plan
    d = load_golden_lsd(path)
    geo = orc.fft_maps(d["nr"], d["inyh"], d["hg"])
    r = orc.rhoofr_lsd(geo, d["c0"], d["f"], d["omega"], d["tpiba2"], d["nsup"])
    assert np.array_equal(r["rhoe"], d["rhoe"]) and r["csums"] == d["csums"]
    p = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], max_batch=2, _cdll=emu_cdll)
    rho, ekin, rg, rr, cs, ca = p.rhoofr_lsd(d["c0"], d["f"], d["nsup"])
    assert np.abs(rho - d["rhoe"]).max() < RTOL * np.abs(d["rhoe"][0]).max()
    assert abs(ekin - d["ekin"]) < ETOL and abs(cs - d["csums"]) < ETOL and abs(ca - d["csumsabs"]) < ETOL
    c2 = d["c2_in"].copy()
    p.vpsi_lsd(d["c0"], c2, d["f"], d["nsup"], d["vpot"])
    assert relmax(c2, d["c2_out"]) < RTOL
    assert len(golden_lsd_cases()) >= 2
