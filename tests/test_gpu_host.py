"""GPU tests of the C++ host (drop-in `dftcxx -i`): whole SCF runs against the reference's own SCF traces stored in the
golden fixtures — total energy within 1e-8 Ha at EQUAL ITERATION INDEX (the reference's 1e-4 stopping rule is loose,
SURVEY.md §8c), same number of iterations, same printed table."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from common import ROOT, TOL_ENERGY, load_golden

from dftcxx_b200 import molecule as M

pytestmark = pytest.mark.gpu

CASES = ["h2o_sto3g", "h2o_p631", "he_sto3g", "co_sto3g_coarse", "h2_sto3g_ultrafine", "ch4_p631_fine", "ethane_p631_fine",
         "benzene_p631_fine", "h2o8_p631_fine"]


def hostlib():
    L = ctypes.CDLL(os.path.join(ROOT, "dftcxx_b200", "libdfthost.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    L.dfthost_scf.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
    L.dfthost_scf2.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp]
    L.dfthost_last_error.restype = ctypes.c_char_p
    return L, dp


MODES = {"device": 0, "host_fused": 1, "host_separate": 2}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", CASES)
def test_scf_energies_match_reference_at_equal_iteration(name, mode):
    """The drop-in host in its three modes: SCF algebra on the device (default), host eigen-solver + fused Fock call on
    page-locked matrices, host eigen-solver + the reference's four grid calls."""
    L, dp = hostlib()
    g = load_golden(name)
    ref = g["scf_energies"]
    nit = len(ref)
    e = np.zeros((nit, 6))
    enuc = ctypes.c_double()
    nb = len(g["bf_nprim"])
    P = np.zeros((nb, nb))
    n = L.dfthost_scf2(os.path.join(M.DATA, "molecules", name + ".in").encode(), 0, 1, MODES[mode], nit, nit, e.ctypes.data_as(dp),
                       ctypes.cast(ctypes.byref(enuc), dp), P.ctypes.data_as(dp))
    assert n == nit, L.dfthost_last_error()
    assert np.array_equal(P, P.T) and abs(2.0 * np.trace(P @ g["scf_S"]) - float(g["nel"])) <= 1e-8  # tr(P S) = nocc
    assert abs(enuc.value - float(g["scf_enuc"])) < 1e-12
    assert np.max(np.abs(e[:, 0] - ref[:, 0])) <= TOL_ENERGY, np.abs(e[:, 0] - ref[:, 0])
    assert np.max(np.abs(e[:, 1] - ref[:, 1])) <= TOL_ENERGY  # E_xc
    assert np.max(np.abs(e[:, 2] - ref[:, 2])) <= TOL_ENERGY  # E_one
    assert np.max(np.abs(e[:, 3] - ref[:, 3])) <= TOL_ENERGY  # E_J
    assert np.max(np.abs(e[:, 4] - ref[:, 4])) <= 1e-9        # electron count


@pytest.mark.parametrize("name", ["h2o_p631", "benzene_p631_fine", "h2o8_p631_fine"])
def test_scf_on_two_gpus_from_one_process(name):
    """`dftcxx -i ... --gpus 2`: the C++ host drives two devices through ONE dftgrid_create_multi handle."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    L, dp = hostlib()
    g = load_golden(name)
    ref = g["scf_energies"]
    nit = len(ref)
    for mode in (0, 1):
        e = np.zeros((nit, 6))
        n = L.dfthost_scf2(os.path.join(M.DATA, "molecules", name + ".in").encode(), 0, 2, mode, nit, nit, e.ctypes.data_as(dp), None, None)
        assert n == nit, L.dfthost_last_error()
        assert np.max(np.abs(e[:, 0] - ref[:, 0])) <= TOL_ENERGY, np.abs(e[:, 0] - ref[:, 0])
        assert np.max(np.abs(e[:, 4] - ref[:, 4])) <= 1e-9
    exe = os.path.join(ROOT, "dftcxx_b200", "bin", "dftcxx")
    r = subprocess.run([exe, "-i", os.path.join(M.DATA, "molecules", name + ".in"), "--gpus", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = re.findall(r"^\s*(\d+)\s+(-?\d+\.\d{7})\s+(\d+\.\d\d) \(\s*(\d+)\)", r.stdout, flags=re.M)
    assert len(rows) >= nit
    for (it, et, nel, nelec), rr in zip(rows, ref):
        assert abs(float(et) - rr[0]) < 1.5e-7


def fermi_conditioning(name, nocc):
    """eps * width / gap of the core-guess F' = X^T H X at the Fermi level (S, T, V from the device): how far the projector onto
    the nocc lowest eigenvectors — and with it P and every energy of the trace — is defined by the matrices at all."""
    from dftcxx_b200.grid import MolecularGrid

    mol = M.Molecule.from_file(os.path.join(M.DATA, "molecules", name + ".in"))
    mg = MolecularGrid(mol)
    mg.set_grid_parameters(10, 4, 5)
    mg.create_grid()
    S, T, V = mg.one_electron()
    mg.close()
    w, U = np.linalg.eigh(S)
    X = U / np.sqrt(w)
    e = np.linalg.eigvalsh(X.T @ (T + V) @ X)
    gap, width = e[nocc] - e[nocc - 1], e[-1] - e[0]
    return np.finfo(float).eps * width / max(gap, 1e-300), gap


@pytest.mark.parametrize("name", ["h2o32_p631_fine", "c40h82_p631_fine"])
def test_large_scf_trace_matches_reference_iteration_by_iteration(name):
    """BASELINE configs 4 and 5 (i): the drop-in host (device integrals, own orthogonalisation, device-resident algebra) against
    the unmodified reference's SCF iterations — total energy within 1e-8 Ha at equal iteration index.  (The reference's 50 %
    mixing does not converge (H2O)32: its energies go -2195.55 -> -2211.83 -> ...; the point is to match it, not to fix it.)

    C40H82 is the exception that the test documents instead of hiding: the core-guess F' of the all-trans chain has a
    near-degenerate pair of orbitals AT the Fermi level (gap 4.7e-10 Ha between orbitals 161 and 162, width 28.7 Ha), so the
    projector onto "the 161 lowest eigenvectors" is defined by the matrices only to eps * width / gap = 1.4e-5 — whichever
    eigen-solver runs (the oracle build's Jacobi shim, real Eigen, the host's QL, the device's purification) picks its own
    rotation inside the pair and the traces part ways at the 1e-6 Ha level in iteration 1.  There the trace is held to
    conditioning * 1 Ha and the test asserts the degeneracy itself; parity at this configuration is pinned by the fixed-P pass
    and by the reference's OWN iteration-2 density matrix (tests/test_gpu_large_parity.py: J, XC 1e-10, energies 1e-8)."""
    from common import have_golden

    if not have_golden(name):
        pytest.skip("fixture not generated yet")
    L, dp = hostlib()
    g = load_golden(name)
    ref = g["scf_energies"]
    nit = len(ref)
    kappa, gap = fermi_conditioning(name, int(round(float(g["nel"]))) // 2)
    print(name, "Fermi-level gap of the core guess %.3e Ha, conditioning eps*width/gap %.2e" % (gap, kappa))
    if name.startswith("c40h82"):
        assert gap < 1e-8 and kappa > 1e-7  # the near-degeneracy is a property of the molecule, not of a solver
    else:
        assert kappa < 1e-12
    # well conditioned: 1e-8 Ha on every component; otherwise conditioning x 1 Ha, growing with the iterations (the next Fock
    # matrices inherit the ambiguity through P); the components move more than their variationally protected sum
    loose = kappa > 1e-9
    tol = np.array([max(TOL_ENERGY, kappa * 100.0 ** it) for it in range(nit)]) if loose else np.full(nit, TOL_ENERGY)
    ctol = 1e3 * tol if loose else tol
    for mode in (0, 1):
        e = np.zeros((nit, 6))
        enuc = ctypes.c_double()
        n = L.dfthost_scf2(os.path.join(M.DATA, "molecules", name + ".in").encode(), 0, 1, mode, nit, nit, e.ctypes.data_as(dp),
                           ctypes.cast(ctypes.byref(enuc), dp), None)
        assert n == nit, L.dfthost_last_error()
        assert abs(enuc.value - float(g["scf_enuc"])) < 1e-9
        print(name, "mode", mode, "dE", np.abs(e[:, 0] - ref[:, 0]))
        assert np.all(np.abs(e[:, 0] - ref[:, 0]) <= tol), np.abs(e[:, 0] - ref[:, 0])
        assert np.all(np.abs(e[:, 1] - ref[:, 1]) <= ctol) and np.all(np.abs(e[:, 3] - ref[:, 3]) <= ctol)
        assert np.max(np.abs(e[:, 4] - ref[:, 4])) <= 1e-8


def test_scf_stopping_rule_and_iteration_count():
    """Free-running SCF: the reference stops h2o/sto3g after 14 iterations at -72.9906070."""
    L, dp = hostlib()
    e = np.zeros((100, 6))
    n = L.dfthost_scf(os.path.join(M.DATA, "molecules", "h2o_sto3g.in").encode(), 0, 0, 100, e.ctypes.data_as(dp), None)
    assert n == 14
    assert round(e[0, 0], 7) == -72.1721582 and round(e[13, 0], 7) == -72.9906070


def test_cli_prints_the_reference_table():
    exe = os.path.join(ROOT, "dftcxx_b200", "bin", "dftcxx")
    r = subprocess.run([exe, "-i", os.path.join(M.DATA, "molecules", "h2o_p631.in")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for needle in ("Reading input file", "Constructing molecular grid", "Number of radial points: 15", "Lebedev order: 7", "Lmax value: 8",
                   "Starting calculation", "Stopping because energy criterion is reached.", "Total elapsed time:"):
        assert needle in out
    rows = re.findall(r"^\s*(\d+)\s+(-?\d+\.\d{7})\s+(\d+\.\d\d) \(\s*(\d+)\)", out, flags=re.M)
    g = load_golden("h2o_p631")
    assert len(rows) == len(g["scf_energies"]) == 17
    for (it, et, nel, nelec), ref in zip(rows, g["scf_energies"]):
        assert abs(float(et) - ref[0]) < 1.5e-7  # 7 printed decimals
        assert nel == "10.00" and nelec == "10"
    assert "E_XC" in out and "E_NUC" in out and "E_ONE" in out and "E_J" in out


def test_two_electron_integral_mode_is_rejected(tmp_path):
    p = tmp_path / "h2.in"
    p.write_text("name = h2\nbasis = sto3g\nunits = angstrom\nhartree_evaluation = two_electron_integrals\n\nsystem:\n2\nH 0 0 -0.367\nH 0 0 0.367\n")
    exe = os.path.join(ROOT, "dftcxx_b200", "bin", "dftcxx")
    r = subprocess.run([exe, "-i", str(p)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "two_electron_integrals" in r.stderr
