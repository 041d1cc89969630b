import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from common import load_golden, system_from_golden, grid_params
from dftcxx_b200.grid import MolecularGrid
for name in ("h2o_sto3g", "benzene_p631_fine"):
    g = load_golden(name)
    mg = MolecularGrid(system_from_golden(g)); mg.set_grid_parameters(*grid_params(g)); mg.create_grid()
    J, XC, exc, nel = mg.iteration(g["P"])
    print(name, float(np.max(np.abs(J - g["J"]))), float(np.max(np.abs(XC - g["XC"]))))
    mg.close()
