"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

#: north_star tolerance: rho(r) and C2 within 1e-11 relative max-norm; energy within 1e-9 Ha
RTOL = 1e-11
ETOL = 1e-9


def relmax(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("omega", "tpiba2", "ekin", "rsum_g", "rsum_r"):
        d[k] = float(d[k])
    for k in ("group", "ngroups"):
        d[k] = int(d[k])
    d["nr"] = tuple(int(v) for v in d["nr"])
    return d


def golden_lsd_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "lsd", "*.npz")))


def load_golden_lsd(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("omega", "tpiba2", "ekin", "rsum_g", "rsum_r", "csums", "csumsabs"):
        d[k] = float(d[k])
    d["nsup"] = int(d["nsup"])
    d["nr"] = tuple(int(v) for v in d["nr"])
    return d


def golden_vofrho_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "vofrho", "*.npz")))


def load_golden_vofrho(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("omega", "tpiba2"):
        d[k] = float(d[k])
    d["nr"] = tuple(int(v) for v in d["nr"])
    return d


def ener_vector(e):
    """dict(eh, ei, ee, eps, vploc) -> the 9 doubles of the C ABI."""
    return np.array([e["eh"].real, e["eh"].imag, e["ei"].real, e["ei"].imag, e["ee"].real, e["ee"].imag,
                     e["eps"].real, e["eps"].imag, e["vploc"]])


def padded_random(geo, rng):
    """random real field on the mesh in the padded (kr3,kr2,kr1) layout, pads zero"""
    n1, n2, n3 = geo.nr
    a = np.zeros((geo.kr[2], geo.kr[1], geo.kr[0]))
    a[:n3, :n2, :n1] = rng.random((n3, n2, n1)) - 0.3
    return a.reshape(-1)


def golden_kpt_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "kpt", "*.npz")))


def load_golden_kpt(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("omega", "tpiba2", "wk", "ekin", "rsum_g"):
        d[k] = float(d[k])
    d["nr"] = tuple(int(v) for v in d["nr"])
    return d


def golden_tau_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "tau", "*.npz")))


def load_golden_tau(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("omega", "tpiba2"):
        d[k] = float(d[k])
    d["nsup"] = int(d["nsup"])
    d["nr"] = tuple(int(v) for v in d["nr"])
    return d


def golden_fsnip_cases():
    """Fixtures computed by the reference's own Fortran statements (tools/make_golden_fsnip.py, oracle/fsnip.py)."""
    return sorted(glob.glob(os.path.join(GOLDEN, "fsnip", "*.npz")))


def load_golden_fsnip(path):
    d = load_golden(path)
    d["tksham"] = bool(d["tksham"])
    return d
