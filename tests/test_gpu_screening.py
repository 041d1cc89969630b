"""GPU tests of the block-sparsity screening of Phi in the two tensor kernels (csrc/kernels_dense.cuh
k_chunk_masks): per 32-point chunk the 32-column blocks whose amplitudes are all <= tau in magnitude are skipped.
 * tau = 0 skips exact zeros only: rho must not change by a bit against no skipping at all, J / XC / F only by the
   summation order of the stream-K shares (the schedule weights change with the map);
 * the default tau = 1e-20 must be invisible at the parity tolerances (J, XC 1e-10; rho 1e-12 relative);
 * an absurd tau = 1e-6 must change the results visibly (the map is really consulted) but boundedly."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(name, tau, developer=True):
    from dftcxx_b200.grid import MolecularGrid
    from dftcxx_b200.systems import WORKLOADS, synthetic_density

    old = os.environ.get("DFTGRID_SCREEN_TAU")
    old_dev = os.environ.get("DFTGRID_DEVELOPER")
    if tau is None:
        os.environ.pop("DFTGRID_SCREEN_TAU", None)
    else:
        os.environ["DFTGRID_SCREEN_TAU"] = repr(tau)
    if developer:
        os.environ["DFTGRID_DEVELOPER"] = "1"  # the library honours its developer switches only with this set
    else:
        os.environ.pop("DFTGRID_DEVELOPER", None)
    try:
        fac, prm = WORKLOADS[name]
        mol = fac()
        mg = MolecularGrid(mol)
        mg.set_grid_parameters(*prm)
        mg.create_grid()
        P = synthetic_density(mol)
        J, XC, exc, nel = mg.iteration(P)
        rho = mg.get_densities()
        F, ej, _, _ = mg.fock(P)
        frac = mg.screen_fraction()
        mg.close()
    finally:
        if old is None:
            os.environ.pop("DFTGRID_SCREEN_TAU", None)
        else:
            os.environ["DFTGRID_SCREEN_TAU"] = old
        if old_dev is None:
            os.environ.pop("DFTGRID_DEVELOPER", None)
        else:
            os.environ["DFTGRID_DEVELOPER"] = old_dev
    return dict(J=J, XC=XC, F=F, rho=rho, exc=exc, nel=nel, ej=ej, frac=frac)


@pytest.mark.parametrize("name", ["h2o8", "h2o32", "c40h82_fine"])
def test_screening_thresholds(name):
    off = run(name, -1.0)
    exact = run(name, 0.0)
    dflt = run(name, None)
    loose = run(name, 1e-6)
    # (the reported fraction is the contraction's; a basis of one 128-tile has only a diagonal tile, which is never masked)
    assert off["frac"] == 1.0 and exact["frac"] <= 1.0 and dflt["frac"] <= exact["frac"] and loose["frac"] <= dflt["frac"]
    assert name == "h2o8" or loose["frac"] < dflt["frac"]
    # exact zeros only: the density kernel has no cross-CTA split, so rho is bit-identical; the contraction's stream-K
    # shares move with the map, so J / XC / F agree to summation-order rounding
    assert np.array_equal(exact["rho"], off["rho"]) and exact["nel"] == off["nel"]
    scale = max(1.0, np.max(np.abs(off["J"])))
    for k in ("J", "XC", "F"):
        assert np.max(np.abs(exact[k] - off[k])) <= 1e-13 * scale, k
    # default threshold: far below the parity tolerances
    assert np.max(np.abs(dflt["rho"] - off["rho"])) <= 1e-16 * np.max(off["rho"])
    big = off["rho"] > 1e-10 * np.max(off["rho"])
    assert np.max(np.abs(dflt["rho"][big] - off["rho"][big]) / off["rho"][big]) <= 1e-13
    for k in ("J", "XC", "F"):
        assert np.max(np.abs(dflt[k] - off[k])) <= 1e-13 * scale, k
    assert abs(dflt["exc"] - off["exc"]) <= 1e-12 * abs(off["exc"]) and abs(dflt["ej"] - off["ej"]) <= 1e-12 * abs(off["ej"])
    # a loose threshold is visible (the map is consulted) and bounded (errors scale with tau)
    dJ = np.max(np.abs(loose["J"] - off["J"]))
    assert 1e-14 < dJ < 1e-3, dJ
    print(name, "work fraction: exact %.3f default %.3f loose %.3f; dJ(loose) %.2e" % (exact["frac"], dflt["frac"], loose["frac"], dJ))


def test_developer_switches_are_ignored_without_the_master_switch():
    """A stray DFTGRID_* variable must not change a production run: without DFTGRID_DEVELOPER the library does not read them."""
    ref = run("h2o32", None)
    off = run("h2o32", -1.0)                       # developer mode: screening off
    stray = run("h2o32", -1.0, developer=False)    # same variable, master switch absent: ignored
    assert off["frac"] == 1.0 and ref["frac"] < 1.0
    assert stray["frac"] == ref["frac"] and np.array_equal(stray["rho"], ref["rho"])
