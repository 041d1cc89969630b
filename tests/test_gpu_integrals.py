"""GPU parity of the one-electron integrals: S, T, V from k_one_electron through
dftgrid_one_electron against the reference's own matrices (golden scf_S and scf_H = T + V, produced by the unmodified
reference's Taketa-Huzinaga-O-ohata code, src/integrals.cpp:43-387) and against the C++ host's independent
McMurchie-Davidson evaluation of T and V separately.  Tolerances: S 1e-13, H 1e-11 absolute (as for the host integrals in
tests/test_cpu.py), T and V 1e-11 against the host."""
import ctypes
import os

import numpy as np
import pytest

from common import ROOT, load_golden

from dftcxx_b200 import molecule as M

pytestmark = pytest.mark.gpu

CASES = ["h2o_sto3g", "h2o_p631", "he_sto3g", "co_sto3g_coarse", "h2_sto3g_ultrafine", "ch4_p631_fine", "ethane_p631_fine",
         "benzene_p631_fine", "h2o8_p631_fine"]


def grid_of(name):
    from dftcxx_b200.grid import MolecularGrid

    mol = M.Molecule.from_file(os.path.join(M.DATA, "molecules", name + ".in"))
    mg = MolecularGrid(mol)
    mg.set_grid_parameters(10, 4, 5)  # the integrals need the basis tables only: the coarse preset
    mg.create_grid()
    return mol, mg


@pytest.mark.parametrize("name", CASES)
def test_one_electron_integrals_match_reference_and_host(name):
    g = load_golden(name)
    mol, mg = grid_of(name)
    S, T, V = mg.one_electron()
    S2, T2, V2 = mg.one_electron()
    mg.close()
    assert np.array_equal(S, S2) and np.array_equal(T, T2) and np.array_equal(V, V2)  # fixed-order sums
    assert np.array_equal(S, S.T) and np.array_equal(T, T.T) and np.array_equal(V, V.T)
    assert np.max(np.abs(S - g["scf_S"])) < 1e-13
    assert np.max(np.abs(T + V - g["scf_H"])) < 1e-11
    L = ctypes.CDLL(os.path.join(ROOT, "dftcxx_b200", "libdfthost.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    L.dfthost_one_electron.argtypes = [ctypes.c_char_p, ctypes.c_int, dp, dp, dp]
    nb = mol.nbf
    Sh, Th, Vh = (np.zeros((nb, nb)) for _ in range(3))
    assert L.dfthost_one_electron(os.path.join(M.DATA, "molecules", name + ".in").encode(), nb, Sh.ctypes.data_as(dp), Th.ctypes.data_as(dp),
                                  Vh.ctypes.data_as(dp)) == nb
    assert np.max(np.abs(S - Sh)) < 1e-13 and np.max(np.abs(T - Th)) < 1e-11 and np.max(np.abs(V - Vh)) < 1e-11


def test_one_electron_integrals_with_d_shells():
    """Cartesian D shells (unreachable through the reference's parser, which stops at Ar): the device against numerical
    quadrature-free identities — S is the Gram matrix of the amplitudes the grid kernel evaluates, so S_ii agrees with the
    grid's sum(w phi_i^2) for a single atom, and the host's McMurchie-Davidson code for the same synthetic system."""
    from dftcxx_b200.grid import MolecularGrid

    # one Sc-like centre with S, P and all six D functions plus a hydrogen-like S neighbour
    xyz = np.array([[0.1, -0.2, 0.3], [0.9, 1.1, 1.7]])
    Z = np.array([3, 1], dtype=np.int32)
    lmn_list = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]
    alphas = [0.8, 0.25]
    bf_nprim, bf_center, alpha, coeff, norm, lmn = [], [], [], [], [], []
    for (l, m, n) in lmn_list:
        bf_nprim.append(2)
        bf_center.append(xyz[0])
        for a, c in zip(alphas, (0.6, 0.5)):
            alpha.append(a)
            coeff.append(c)
            norm.append(M.gto_norm(a, l, m, n))
            lmn.append((l, m, n))
    bf_nprim.append(1)
    bf_center.append(xyz[1])
    alpha.append(0.5)
    coeff.append(1.0)
    norm.append(M.gto_norm(0.5, 0, 0, 0))
    lmn.append((0, 0, 0))
    sysd = dict(Z=Z, xyz=xyz, bf_nprim=np.array(bf_nprim, np.int32), bf_center=np.array(bf_center), alpha=np.array(alpha), coeff=np.array(coeff),
                norm=np.array(norm), lmn=np.array(lmn, np.int32))
    mg = MolecularGrid(sysd)
    mg.set_grid_parameters(60, 10, 4)
    mg.create_grid()
    S, T, V = mg.one_electron()
    phi, w = mg.get_amplitudes(), mg.get_weights()
    mg.close()
    Sq = phi.T @ (phi * w[:, None])  # quadrature of the same amplitudes on the Becke grid
    assert np.max(np.abs(S - Sq)) < 1e-5
    # kinetic energy of a normalised primitive s Gaussian is 3 alpha / 2; of the hydrogen-like function here: 0.75
    assert abs(T[-1, -1] - 0.75 * S[-1, -1]) < 1e-13
    # nuclear attraction is negative definite on this basis
    assert np.all(np.linalg.eigvalsh(V) < 0.0)
