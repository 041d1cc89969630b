"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tolerances stated by BASELINE.json's north_star:
TOL_MATRIX_ABS = 1e-10  # Kxc / J elementwise, absolute
TOL_REL = 1e-12         # Becke weights and rho, relative
TOL_ENERGY = 1e-8       # total energy, Hartree
# The Becke cell function 0.5*(1 - f3(mu)) cancels catastrophically as mu -> 1, so a tiny weight
# changes by far more than 1e-12 *relative* under any 1-ulp change upstream (glibc pow vs any other cube).  Relative
# error is therefore measured against max(|ref|, floor) with floor = 1e-3 for the dimensionless Becke weight
# (absolute 1e-15) and 1e-6 of the largest value for weights / densities.
BECKE_FLOOR = 1e-3


def load_golden(name):
    """Fixture as a dict.  The large-configuration fixtures (make_golden_large.py) store symmetric matrices as packed upper
    triangles (`X_triu`) and regenerate the synthetic P from its seed: both are expanded here (`X`, `P`)."""
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    nb = len(g["bf_nprim"])
    for k in [k for k in g if k.endswith("_triu")]:
        M = np.zeros((nb, nb))
        M[np.triu_indices(nb)] = g[k]
        g[k[:-5]] = M + np.triu(M, 1).T
    if "P" not in g and "P_checksum" in g:
        from dftcxx_b200.systems import synthetic_density

        class _M:
            nbf = nb
            nelec = int(round(float(g["nel"])))

        P = synthetic_density(_M)
        chk = np.array([P.sum(), np.abs(P).sum(), np.trace(P)])
        assert np.allclose(chk, g["P_checksum"], rtol=1e-13, atol=0), "synthetic P does not reproduce the fixture's"
        g["P"] = P
    return g


def have_golden(name):
    return os.path.exists(os.path.join(GOLDEN, name + ".npz"))


def system_from_golden(g):
    return {k: g[k] for k in ("Z", "xyz", "bf_nprim", "bf_center", "alpha", "coeff", "norm", "lmn")}


def grid_params(g):
    return int(g["radial_points"]), int(g["lebedev_order"]), int(g["lmax"])


def relerr(a, b, floor):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), floor)))
