"""The committed golden vectors (tools/make_golden.py) against both oracle restatements."""
import numpy as np
import pytest

from helpers import ETOL, RTOL, golden_cases, load_golden, relmax
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
