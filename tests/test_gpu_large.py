"""GPU tests at BASELINE.json's large configurations, where no CPU oracle finishes in seconds: size-independent
properties of the path — the dense contractions checked against numpy on the GPU's own Phi / weights / potential,
electron-count normalisation, symmetry, bit-for-bit determinism, invariance of J and XC under P -> 2P."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(name):
    from dftcxx_b200.grid import MolecularGrid
    from dftcxx_b200.systems import WORKLOADS, synthetic_density

    fac, prm = WORKLOADS[name]
    mol = fac()
    g = MolecularGrid(mol)
    g.set_grid_parameters(*prm)
    g.create_grid()
    return mol, g, synthetic_density(mol)


@pytest.mark.parametrize("name", ["h2o32", "c40h82_fine"])
def test_large_configuration_properties(name):
    mol, g, P = build(name)
    try:
        J, XC, exc, nel = g.iteration(P)
        assert abs(nel - mol.nelec) <= 1e-8
        assert np.array_equal(J, J.T) and np.array_equal(XC, XC.T)
        J2, XC2, exc2, nel2 = g.iteration(P)
        assert np.array_equal(J, J2) and np.array_equal(XC, XC2) and exc == exc2  # deterministic reductions
        # dense contractions against numpy on the device's own operands (nb = 416 / 524: several tile pairs, a ragged
        # edge tile, diagonal tiles, dozens of stream-K segments)
        phi = g.get_amplitudes()
        w = g.get_weights()
        V = g.get_potential()
        rho = g.get_densities()
        Jn = 0.5 * (phi.T * (w * V)) @ phi
        assert np.max(np.abs(J - Jn)) <= 1e-10 * max(1.0, np.max(np.abs(Jn)))
        raw = 2.0 * np.einsum("pi,pi->p", phi @ P, phi)
        scale = mol.nelec / np.dot(w, raw)
        assert np.max(np.abs(rho - raw * scale)) <= 1e-12 * np.max(rho)
        assert abs(np.dot(w, rho) - mol.nelec) <= 1e-8
        # E_J = 2 tr(P J) equals the pointwise sum 0.5 * sum w V rho_unscaled
        assert abs(2.0 * np.trace(P @ J) - 0.5 * np.dot(w * V, raw)) <= 1e-8 * abs(np.trace(P @ J))
        # the Hartree potential is positive and decays: a sanity anchor for the multipole / spline machinery
        assert V.min() > 0.0
        J3, XC3, exc3, nel3 = g.iteration(2.0 * P)
        assert np.max(np.abs(J3 - J)) <= 1e-9 and np.max(np.abs(XC3 - XC)) <= 1e-9 and abs(exc3 - exc) <= 1e-9
    finally:
        g.close()


def test_north_star_workload_runs_and_normalises():
    """(H2O)64 / 6-31G / fine: 560 640 points x 832 basis functions on one GPU."""
    mol, g, P = build("h2o64")
    try:
        J, XC, exc, nel = g.iteration(P)
        assert (g.npoints, mol.nbf) == (560640, 832)
        assert abs(nel - 640.0) <= 1e-8
        assert np.array_equal(J, J.T) and np.array_equal(XC, XC.T)
        assert np.all(np.isfinite(J)) and np.all(np.isfinite(XC)) and exc < 0.0
        # XC is negative definite-ish on the diagonal (v_xc < 0), J positive
        assert np.all(np.diag(XC) < 0.0) and np.all(np.diag(J) > 0.0)
        t = g.timings()
        assert t["total"] < 500.0  # ms; a CPU fallback would take minutes
    finally:
        g.close()
