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
