import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.refpy import Ref, DATA
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.molecule import Molecule
from dftcxx_b200.systems import synthetic_density
name = sys.argv[1]
path = os.path.join(DATA, "molecules", name)
r = Ref(path); mol = Molecule.from_file(path)
g = MolecularGrid(mol); g.set_grid_parameters(r.nrad, r.lebedev_order, r.lmax); g.create_grid()
P = synthetic_density(mol)
r.set_density(P); Jr = r.hartree(); hi = r.hartree_intermediates()
g.set_density(P); Jg = g.calculate_hartree_potential()
V = g.get_potential(); xyz = g.get_positions()
d = np.abs(V - hi["V"])
order = np.argsort(-d)[:12]
npa = r.nrad * r.nang
for i in order:
    own = i // npa; sh = (i % npa) // r.nang; a = i % r.nang
    rel = xyz[i] - mol.xyz
    print(i, "own", own, "shell", sh, "ang", a, "dV %.3e" % d[i], "V %.6f" % V[i], "xyz", xyz[i], "dist", np.linalg.norm(rel, axis=1),
          "sin(theta) to atoms", np.hypot(rel[:, 0], rel[:, 1]) / np.linalg.norm(rel, axis=1))
print("Vfuzzy diff", np.max(np.abs(hi["V_fuzzy"] - 0)), "U diff", np.max(np.abs(g.get_U_lm() - hi["U_lm"])))
x = np.array(sorted((1 + np.cos(np.pi * np.arange(1, r.nrad + 1) / (r.nrad + 1))) / (1 - np.cos(np.pi * np.arange(1, r.nrad + 1) / (r.nrad + 1)))))
print("nodes", x)
