import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import *
from dftcxx_b200.grid import MolecularGrid
g = load_golden("ch4_p631_dense422")
mg = MolecularGrid(system_from_golden(g)); mg.set_grid_parameters(*grid_params(g)); mg.create_grid()
J, XC, exc, nel = mg.iteration(g["P"])
ri = g["rad_idx"]
U = mg.get_U_lm()[:, ri]; dU = np.abs(U - g["U_lm"])
print("U max abs", dU.max(), "at", np.unravel_index(dU.argmax(), dU.shape), "scale", np.abs(g["U_lm"]).max())
for l in range(0, 12):
    sl = slice(l * l, (l + 1) ** 2)
    print(" l=%d  max|dU| %.3e  max|U| %.3e" % (l, dU[:, :, sl].max(), np.abs(g["U_lm"][:, :, sl]).max()))
print("J max abs", np.max(np.abs(J - g["J"])), "scale", np.abs(g["J"]).max())
V = mg.get_potential()[g["idx"]]; print("V max abs", np.max(np.abs(V - g["V"])), "scale", np.abs(g["V"]).max())
print("rho_lm max abs", np.max(np.abs(mg.get_rho_lm()[:, ri] - g["rho_lm"])), np.abs(g["rho_lm"]).max())
print(mg.timings())
