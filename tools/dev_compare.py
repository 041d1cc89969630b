"""Developer script: run one molecule through the CUDA library and the reference oracle and print error norms."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.refpy import Ref, DATA
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.molecule import Molecule

def rel(a, b, floor=0.0):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if floor > 0 else float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

def main(name):
    path = os.path.join(DATA, "molecules", name)
    t = time.time(); r = Ref(path, full=False); print("ref grid build %.2fs" % (time.time() - t), r.natoms, r.nbf, r.npts)
    sysd = r.system()
    mol = Molecule.from_file(path)
    for k in ("Z", "xyz", "bf_nprim", "bf_center", "alpha", "coeff", "norm", "lmn"):
        same = np.array_equal(np.asarray(sysd[k]), getattr(mol, k))
        if not same: print("  parser mismatch in", k, np.max(np.abs(np.asarray(sysd[k], float) - getattr(mol, k))))
    g = MolecularGrid(mol); g.set_grid_parameters(r.nrad, r.lebedev_order, r.lmax)
    t = time.time(); g.create_grid(); print("gpu create_grid %.3fs" % (time.time() - t), g.timings())
    xyz, w, wb = r.grid()
    gx, gw, gwb = g.get_positions(), g.get_weights(), g.get_becke_weights()
    print("pos max abs diff", np.max(np.abs(gx - xyz)), "bitwise equal:", np.array_equal(gx, xyz))
    print("becke: bitwise equal frac %.6f  max abs %.3e  max rel(floor 1e-3) %.3e" % (np.mean(gwb == wb), np.max(np.abs(gwb - wb)), rel(gwb, wb, 1e-3)))
    print("weights: max rel(floor=1e-3*max) %.3e" % rel(gw, w, 1e-3 * np.max(np.abs(w))))
    phi = r.amplitudes(); gphi = g.get_amplitudes()
    print("phi: max abs %.3e  max rel(floor 1e-10) %.3e  bitwise frac %.4f" % (np.max(np.abs(gphi - phi)), rel(gphi, phi, 1e-10), np.mean(gphi == phi)))
    rng = np.random.default_rng(1)
    nocc = max(1, r.nelec // 2)
    Cm = rng.standard_normal((r.nbf, nocc)) / np.sqrt(r.nbf)
    P = Cm @ Cm.T
    raw = r.set_density_raw(P); rho_raw = r.densities()
    r.set_density(P); rho = r.densities()
    g.set_density(P); grho = g.get_densities()
    print("rho: max rel %.3e (floor 1e-12*max: %.3e)  nel ref %.12f gpu %.12f" % (rel(grho, rho, 1e-300), rel(grho, rho, 1e-12 * rho.max()), r.electron_count(), g.calculate_density()))
    t = time.time(); Jr = r.hartree(); print("ref hartree %.2fs" % (time.time() - t))
    t = time.time(); Jg = g.calculate_hartree_potential(); print("gpu hartree %.4fs" % (time.time() - t))
    hi = r.hartree_intermediates()
    print("rho_lm max abs %.3e (scale %.3e)" % (np.max(np.abs(g.get_rho_lm() - hi["rho_lm"])), np.max(np.abs(hi["rho_lm"]))))
    print("U_lm   max abs %.3e (scale %.3e)" % (np.max(np.abs(g.get_U_lm() - hi["U_lm"])), np.max(np.abs(hi["U_lm"]))))
    print("V      max abs %.3e (scale %.3e)" % (np.max(np.abs(g.get_potential() - hi["V"])), np.max(np.abs(hi["V"]))))
    print("J      max abs %.3e (scale %.3e)" % (np.max(np.abs(Jg - Jr)), np.max(np.abs(Jr))))
    XCr, excr = r.xc(); XCg, excg = g.calculate_exchange_correlation()
    print("XC     max abs %.3e (scale %.3e)  exc ref %.12f gpu %.12f diff %.3e" % (np.max(np.abs(XCg - XCr)), np.max(np.abs(XCr)), excr, excg, excg - excr))
    J2, XC2, exc2, nel2 = g.iteration(P)
    print("iteration(): J diff vs 4-call %.3e XC diff %.3e" % (np.max(np.abs(J2 - Jg)), np.max(np.abs(XC2 - XCg))), g.timings())

if __name__ == "__main__":
    for n in sys.argv[1:] or ["h2o_sto3g.in"]:
        print("=====", n); main(n)
