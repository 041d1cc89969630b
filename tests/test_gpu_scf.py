"""GPU tests of the device-resident SCF algebra (dftgrid_scf_*: F' = X^T F X, purified projector, P = X D' X^T, mixing,
E_one) against the reference's own SCF traces in the golden fixtures: total energy within 1e-8 Ha AT EQUAL ITERATION INDEX,
the density matrix of the probe iteration elementwise, and the purified projector against numpy's eigenvectors of the
device's own F'.  H and X are the reference's matrices (fixture), so only the device algebra + grid path are under test."""
import numpy as np
import pytest

from common import TOL_ENERGY, grid_params, load_golden, system_from_golden

pytestmark = pytest.mark.gpu

CASES = ["h2o_sto3g", "h2o_p631", "he_sto3g", "co_sto3g_coarse", "h2_sto3g_ultrafine", "ch4_p631_fine", "ethane_p631_fine",
         "benzene_p631_fine", "h2o8_p631_fine"]


def make_grid(g, **kw):
    from dftcxx_b200.grid import MolecularGrid

    mg = MolecularGrid(system_from_golden(g), **kw)
    mg.set_grid_parameters(*grid_params(g))
    mg.create_grid()
    return mg


def run_trace(mg, g, nit):
    nocc = int(round(float(g["nel"]))) // 2
    mg.scf_init(g["scf_H"], g["scf_X"], nocc, 0.5)
    first = mg.scf_step(include_xc=False)  # DFT::construct_matrices: core guess, J(P0) only
    rows, probes = [], {}
    for it in range(1, nit + 1):
        r = mg.scf_step()
        rows.append([r["e_one"] + r["e_j"] + float(g["scf_enuc"]) + r["exc"], r["exc"], r["e_one"], r["e_j"], r["nel"], r["purification_steps"],
                     r["idempotency"]])
        if it == int(g["scf_probe_iter"]):
            probes["P"] = mg.scf_matrix("P")
            probes["Fp"] = mg.scf_matrix("F_prime")
            probes["D"] = mg.scf_matrix("D_prime")
    return first, np.array(rows), probes


@pytest.mark.parametrize("name", CASES)
def test_device_scf_matches_reference_trace(name):
    g = load_golden(name)
    ref = g["scf_energies"]
    mg = make_grid(g)
    try:
        first, e, pr = run_trace(mg, g, len(ref))
        assert np.max(np.abs(e[:, 0] - ref[:, 0])) <= TOL_ENERGY, np.abs(e[:, 0] - ref[:, 0])
        for k in (1, 2, 3):  # E_xc, E_one, E_J
            assert np.max(np.abs(e[:, k] - ref[:, k])) <= TOL_ENERGY
        assert np.max(np.abs(e[:, 4] - ref[:, 4])) <= 1e-9
        # the density matrix of the probe iteration, elementwise: P = X D' X^T, so an error of D' (conditioned like
        # eps*width/gap in the reference's eigenvectors as much as here) is amplified by |X|_2^2 = 1/lambda_min(S)
        # (538 for the toy ethane geometry, 1435 for benzene): 1e-10, or 1e-12 |X|_2^2 where that is larger
        assert np.max(np.abs(pr["P"] - g["scf_P"])) <= max(1e-10, 1e-12 * np.linalg.norm(g["scf_X"], 2) ** 2)
        assert np.array_equal(pr["P"], pr["P"].T)
        # the purified projector against the eigenvectors of the device's own F'
        n = pr["Fp"].shape[0]
        nocc = int(round(float(g["nel"]))) // 2
        w, C = np.linalg.eigh(pr["Fp"])
        if 0 < nocc < n:
            gap = w[nocc] - w[nocc - 1]
            De = C[:, :nocc] @ C[:, :nocc].T
            assert np.max(np.abs(pr["D"] - De)) <= max(1e-12, 1e-14 * (w[-1] - w[0]) / gap)
            assert abs(np.trace(pr["D"]) - nocc) <= 1e-11
            assert np.max(np.abs(pr["D"] @ pr["D"] - pr["D"])) <= 1e-11
        assert np.all(e[:, 5] <= 64)  # purification steps
    finally:
        mg.close()


def test_device_scf_is_deterministic_and_multi_gpu_identical():
    import torch

    g = load_golden("benzene_p631_fine")
    mg = make_grid(g)
    try:
        _, e1, p1 = run_trace(mg, g, 3)
    finally:
        mg.close()
    mg = make_grid(g)
    try:
        _, e2, p2 = run_trace(mg, g, 3)
    finally:
        mg.close()
    assert np.array_equal(e1, e2) and np.array_equal(p1["P"], p2["P"])  # bit for bit, run to run
    if torch.cuda.device_count() >= 2:
        mg = make_grid(g, ngpus=2)
        try:
            _, e3, p3 = run_trace(mg, g, 3)
        finally:
            mg.close()
        # two devices sum the shards' F in a different order than one device: rounding-level differences, amplified in P by
        # |X|_2^2 (1435 for benzene) like in the fixture comparison above
        assert np.max(np.abs(e3[:, :5] - e1[:, :5])) <= 1e-9
        assert np.max(np.abs(p3["P"] - p1["P"])) <= max(1e-11, 1e-13 * np.linalg.norm(g["scf_X"], 2) ** 2)
