"""Synthetic workloads of BASELINE.json / SURVEY 8d: random, normalised, decaying wavefunctions,
occupations f=2, V(r) ~ U(-1,0) on mesh points and 0 on the pads.  Deterministic in (n, nstate).
"""
from __future__ import annotations

import numpy as np

from .api import leadim
from .gvec import dotp_weights, half_sphere


def make_inputs(n, nstate, seed=None, f_pattern="all2", out_c0=None):
    """Returns dict(nr, kr, inyh, hg, c0 (nstate, ngw) complex128, f, vpot (nnr1,), tpiba2, omega)."""
    nr = (n, n, n) if isinstance(n, int) else tuple(n)
    inyh, hg = half_sphere(nr)
    ngw = hg.shape[0]
    if seed is None:
        seed = 1234 + nr[0] + 7 * nstate
    rng = np.random.default_rng(seed)
    gcutw = (min(nr) / 4.0) ** 2
    damp = np.exp(-hg / (0.25 * gcutw))
    w = dotp_weights(ngw, True)
    c0 = out_c0 if out_c0 is not None else np.empty((nstate, ngw), dtype=np.complex128)
    for i in range(nstate):
        re = rng.standard_normal(ngw)
        im = rng.standard_normal(ngw)
        c = (re + 1j * im) * damp
        c[0] = c[0].real
        nrm = float(np.dot(w, c.real ** 2 + c.imag ** 2)) - 0.0
        # dotp counts Re^2 only for G=0; Im(c[0]) is 0 so the expression above is exact
        c0[i] = c / np.sqrt(nrm)
    f = np.full(nstate, 2.0)
    if f_pattern == "mixed":
        f[::3] = 1.0
        f[1::5] = 0.0
    kr = tuple(leadim(v) for v in nr)
    v = np.zeros((kr[2], kr[1], kr[0]))
    v[:nr[2], :nr[1], :nr[0]] = -rng.random((nr[2], nr[1], nr[0]))
    return dict(nr=nr, kr=kr, inyh=inyh, hg=hg, c0=c0, f=f, vpot=v.reshape(-1), tpiba2=1.0, omega=1.0)
