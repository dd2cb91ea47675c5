"""Device-side parity check made of REFERENCE code: the reference's own GPU path for rhoofr / vpsi - its
CUDA sources (src/cuuser_utils.cu, src/cuuser_utils_kernels.cu) compiled by nvcc where they lie, cuFFT in
the role of mltfft_cuda, the stage order of fftcu_methods.mod.F90 incl. its host round trip per transform
(oracle/ref_gpu_driver.cu) - against this library's kernels through the C ABI, and against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cpmd_b200 import Plan  # noqa: E402
from helpers import RTOL, relmax  # noqa: E402
from oracle import cpmd_oracle as orc  # noqa: E402
from oracle import ref_gpu  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    if ref_gpu.load() is None:
        pytest.skip("oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)")
    return torch.device("cuda:0")


@pytest.mark.parametrize("nr,ns,scatter", [(16, 5, False), ((16, 20, 24), 4, True), (48, 7, False), (96, 4, True),
                                           (192, 4, False)])
def test_library_matches_the_reference_gpu_path(dev, nr, ns, scatter):
    geo = orc.make_geometry(nr)
    tpiba2, omega = 0.9, 1.3
    c0, f, v = orc.synthetic_inputs(geo, ns, f_pattern="mixed")
    ref = ref_gpu.RefGpu(geo, tpiba2, omega)
    assert ref.L.refgpu_source().decode().endswith("reference/src")
    rho_ref = ref.rhoofr(c0, f, device_scatter=scatter)
    c2_ref = ref.vpsi(c0, 0.5 * c0, f, v, device_scatter=scatter)
    assert ref.launches > 0
    plan = Plan(geo.nr, geo.inyh, geo.hg, tpiba2, omega, max_batch=2)
    c0d = torch.from_numpy(c0).to(dev)
    rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    plan.rhoofr_dev(c0d, f, rho)
    c2 = 0.5 * c0d
    plan.vpsi_dev(c0d, c2, f, torch.from_numpy(v).to(dev))
    assert relmax(rho.cpu().numpy(), rho_ref) < RTOL
    assert relmax(c2.cpu().numpy(), c2_ref) < RTOL
    if max(geo.nr) <= 96:       # and the oracle agrees with the reference's GPU path
        assert relmax(orc.rhoofr(geo, c0, f, omega, tpiba2)["rhoe"], rho_ref) < RTOL
        assert relmax(orc.vpsi(geo, c0, 0.5 * c0, f, v, tpiba2), c2_ref) < RTOL
    ref.close()
