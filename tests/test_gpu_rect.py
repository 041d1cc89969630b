"""GPU parity of the density dump: density and density gradient on the reference's RectangularGrid
(src/rectangulargrid.cpp:34-95) through dftgrid_rectangular_density, against the fixture the unmodified reference produced
(tests/golden/make_golden_rect.py), and the C++ host's DFT::finalize dump file (`density_dump = <file>` in the input).
Tolerances: positions bit-identical; rho 1e-12 relative (floor 1e-6 of the maximum, like the grid density); gradient
1e-12 of its largest component."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import GOLDEN, ROOT, TOL_REL, relerr

from dftcxx_b200 import molecule as M

pytestmark = pytest.mark.gpu

CASES = ["h2o_sto3g", "benzene_p631_fine", "co_sto3g_coarse"]


def fixture():
    return np.load(os.path.join(GOLDEN, "rect_density.npz"))


def engine(z, name):
    from dftcxx_b200.grid import MolecularGrid

    sysd = {k: z["%s.%s" % (name, k)] for k in ("Z", "xyz", "bf_nprim", "bf_center", "alpha", "coeff", "norm", "lmn")}
    mg = MolecularGrid(sysd)
    mg.set_grid_parameters(int(z[name + ".radial_points"]), int(z[name + ".lebedev_order"]), int(z[name + ".lmax"]))
    mg.create_grid()
    return mg


@pytest.mark.parametrize("name", CASES)
def test_rectangular_density_and_gradient_match_reference(name):
    z = fixture()
    mg = engine(z, name)
    size, dp = float(z[name + ".size"]), int(z[name + ".dp"])
    pos, rho, grad = mg.rectangular_density(size, dp, z[name + ".P"])
    assert np.array_equal(pos, z[name + ".pos"]), "box points must match the reference bit for bit"
    assert relerr(rho, z[name + ".rho"], 1e-6 * np.max(np.abs(z[name + ".rho"]))) <= TOL_REL
    gmax = np.max(np.abs(z[name + ".grad"]))
    assert np.max(np.abs(grad - z[name + ".grad"])) <= 1e-12 * gmax
    # twice the same call: bit-identical (fixed-order reductions)
    pos2, rho2, grad2 = mg.rectangular_density(size, dp, z[name + ".P"])
    assert np.array_equal(rho, rho2) and np.array_equal(grad, grad2)
    if name + ".P_nonsym" in z:
        # GridPoint::set_gradient keeps both product-rule terms, so a non-symmetric P is well defined too
        _, rho_n, grad_n = mg.rectangular_density(size, dp, z[name + ".P_nonsym"])
        assert relerr(rho_n, z[name + ".rho_nonsym"], 1e-6 * np.max(np.abs(z[name + ".rho_nonsym"]))) <= 1e-11
        assert np.max(np.abs(grad_n - z[name + ".grad_nonsym"])) <= 1e-11 * np.max(np.abs(z[name + ".grad_nonsym"]))
    mg.close()


def test_rectangular_density_rejects_bad_arguments():
    from dftcxx_b200.grid import GridError

    z = fixture()
    mg = engine(z, "h2o_sto3g")
    with pytest.raises(GridError):
        mg.rectangular_density(5.0, 1, z["h2o_sto3g.P"])
    with pytest.raises(GridError):
        mg.rectangular_density(-1.0, 5, z["h2o_sto3g.P"])
    mg.close()


def test_host_finalize_writes_the_reference_dump_format(tmp_path):
    """`dftcxx -i` with `density_dump = <file>`: DFT::finalize (src/dft.cpp:489-504) writes x y z grad_x grad_y grad_z in the
    reference's "%12.8f  " format for the converged density; the numbers are the engine's for that P, and the text layout
    is the one RectangularGrid::write_gradient of the reference produced (fixture lines)."""
    z = fixture()
    L = ctypes.CDLL(os.path.join(ROOT, "dftcxx_b200", "libdfthost.so"))
    dp_t = ctypes.POINTER(ctypes.c_double)
    L.dfthost_scf2.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp_t, dp_t, dp_t]
    L.dfthost_last_error.restype = ctypes.c_char_p
    src = open(os.path.join(M.DATA, "molecules", "h2o_sto3g.in")).read()
    dump = tmp_path / "data.dat"
    inp = tmp_path / "h2o_dump.in"
    inp.write_text("density_dump = %s\ndensity_dump_size = 5.0\ndensity_dump_points = 15\n%s" % (dump, src))
    e = np.zeros((100, 6))
    P = np.zeros((7, 7))
    n = L.dfthost_scf2(str(inp).encode(), 0, 1, -1, 0, 100, e.ctypes.data_as(dp_t), None, P.ctypes.data_as(dp_t))
    assert n > 3, L.dfthost_last_error()
    lines = dump.read_text().splitlines()
    assert len(lines) == 15 ** 3
    ref_lines = [str(s) for s in z["h2o_sto3g.dump_lines"]]
    pat = re.compile(r"^( *-?\d+\.\d{8}  ){5} *-?\d+\.\d{8}$")
    assert all(pat.match(s) for s in lines) and all(pat.match(s) for s in ref_lines)
    assert [len(s) for s in lines[:40]] == [len(s) for s in ref_lines[:40]]
    assert [s[:40] for s in lines[:40]] == [s[:40] for s in ref_lines[:40]]  # the three position columns, verbatim
    vals = np.array([[float(t) for t in s.split()] for s in lines])
    mg = engine(z, "h2o_sto3g")
    pos, rho, grad = mg.rectangular_density(5.0, 15, P)
    mg.close()
    assert np.max(np.abs(vals[:, :3] - pos)) <= 5.1e-9 and np.max(np.abs(vals[:, 3:] - grad)) <= 5.1e-9
