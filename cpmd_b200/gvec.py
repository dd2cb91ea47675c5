"""Host-side G-vector generation for synthetic workloads (mirror of what ``loadpa``/``rggen``
hand to the hot path: ``inyh(3,ngw)`` and ``hg(ngw)``; loadpa_utils.mod.F90:282-335,
rggen_utils.mod.F90:121-129).  In a real CPMD run these arrays come from the host program; the
library only needs them as inputs to :class:`cpmd_b200.api.Plan`.
"""
from __future__ import annotations

import numpy as np


def half_sphere(nr, gcutw=None):
    """Plane waves of the wavefunction cutoff for an orthorhombic cell in units where
    b_d = e_d: all integer (i,j,k) with i^2+j^2+k^2 < gcutw in the half space
    (i>0) or (i==0, j>0) or (i==j==0, k>=0), sorted by |G|^2 with G=0 first.

    Returns (inyh int32 (3,ngw) 1-based, hg float64 (ngw,)).  gcutw defaults to (min(nr)/4)^2
    (dual = 4)."""
    if isinstance(nr, int):
        nr = (nr, nr, nr)
    if gcutw is None:
        gcutw = (min(nr) / 4.0) ** 2
    m = int(np.floor(np.sqrt(gcutw))) + 1
    ax = np.arange(-m, m + 1, dtype=np.int64)
    gi, gj, gk = np.meshgrid(np.arange(0, m + 1, dtype=np.int64), ax, ax, indexing="ij")
    gi, gj, gk = gi.reshape(-1), gj.reshape(-1), gk.reshape(-1)
    g2 = gi * gi + gj * gj + gk * gk
    half = (gi > 0) | ((gi == 0) & ((gj > 0) | ((gj == 0) & (gk >= 0))))
    keep = half & (g2 < gcutw)
    gi, gj, gk, g2 = gi[keep], gj[keep], gk[keep], g2[keep]
    order = np.argsort(g2, kind="stable")
    gi, gj, gk, g2 = gi[order], gj[order], gk[order], g2[order]
    nh = [n // 2 + 1 for n in nr]
    inyh = np.stack([gi + nh[0], gj + nh[1], gk + nh[2]]).astype(np.int32)
    return inyh, g2.astype(np.float64)


def dotp_weights(ngw, geq0=True):
    """Weights of dotp (dotp_utils.mod.F90:26-53): 2 on the half sphere, 1 for G=0."""
    w = np.full(ngw, 2.0)
    if geq0:
        w[0] = 1.0
    return w
