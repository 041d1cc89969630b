"""Developer script: one-time setup costs on a workload — grid build, one-electron integrals on the device vs the C++ host,
density dump.  usage: dev_setup_times.py <workload>"""
import ctypes
import os

os.environ.setdefault("DFTGRID_DEVELOPER", "1")  # developer script: the library's A/B switches are live
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dftcxx_b200 import molecule as M
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import synthetic_density

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(name):
    path = os.path.join(M.DATA, "molecules", name + ".in")
    mol = M.Molecule.from_file(path)
    st = mol.settings
    g = MolecularGrid(mol)
    g.set_grid_parameters(st.radial_points, st.lebedev_order, st.lmax)
    t = time.time()
    g.create_grid()
    print(name, "natoms", mol.natoms, "nbf", mol.nbf, "create_grid wall %.3f s" % (time.time() - t), flush=True)
    for _ in range(3):
        t = time.time()
        S, T, V = g.one_electron()
        print("  device one_electron wall %.1f ms" % ((time.time() - t) * 1e3), flush=True)
    L = ctypes.CDLL(os.path.join(ROOT, "dftcxx_b200", "libdfthost.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    L.dfthost_one_electron.argtypes = [ctypes.c_char_p, ctypes.c_int, dp, dp, dp]
    nb = mol.nbf
    Sh, Th, Vh = (np.zeros((nb, nb)) for _ in range(3))
    t = time.time()
    L.dfthost_one_electron(path.encode(), nb, Sh.ctypes.data_as(dp), Th.ctypes.data_as(dp), Vh.ctypes.data_as(dp))
    print("  host one_electron (serial entry point) wall %.1f ms; max|dS| %.2e max|dT| %.2e max|dV| %.2e" % (
        (time.time() - t) * 1e3, np.max(np.abs(S - Sh)), np.max(np.abs(T - Th)), np.max(np.abs(V - Vh))), flush=True)
    P = synthetic_density(mol)
    for dp_ in (15, 15, 41, 41):
        t = time.time()
        pos, rho, grad = g.rectangular_density(5.0, dp_, P)
        print("  rectangular_density %d^3 wall %.1f ms" % (dp_, (time.time() - t) * 1e3), flush=True)
    g.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "h2o8_p631_fine")
