"""Developer script: a few device-resident SCF steps of a workload with a synthetic H / X (for ncu captures of k_gemm_nn / k_pm_*)."""
import os

os.environ.setdefault("DFTGRID_DEVELOPER", "1")  # developer script: the library's A/B switches are live
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "h2o64"
fac, prm = WORKLOADS[name]
mol = fac()
g = MolecularGrid(mol)
g.set_grid_parameters(*prm)
g.create_grid()
nb = mol.nbf
rng = np.random.default_rng(1)
A = rng.standard_normal((nb, nb))
H = -(A @ A.T) / nb - np.diag(np.linspace(0, 20, nb)[::-1])
Q, _ = np.linalg.qr(rng.standard_normal((nb, nb)))
g.scf_init(H, Q, mol.nelec // 2, 0.5)
print(g.scf_step(include_xc=False))
for _ in range(2):
    print(g.scf_step())
g.close()
