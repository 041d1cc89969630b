"""Developer script (round 2): per-phase device timings of the two-matrix iteration, the fused Fock build and the
device-resident SCF step for a synthetic workload.  usage: dev_perf2.py <workload> [ngpus]"""
import os

os.environ.setdefault("DFTGRID_DEVELOPER", "1")  # developer script: the library's A/B switches are live
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import WORKLOADS, synthetic_density


def main(name, ngpus=1, iters=4):
    fac, prm = WORKLOADS[name]
    mol = fac()
    g = MolecularGrid(mol, ngpus=ngpus)
    g.set_grid_parameters(*prm)
    t = time.time()
    g.create_grid()
    print(name, "ngpus", ngpus, "natoms", mol.natoms, "nbf", mol.nbf, "npts", g.npoints, "create_grid wall %.3fs" % (time.time() - t),
          {k: round(v, 3) for k, v in g.timings().items() if k in ("points", "becke", "phi")}, flush=True)
    P = synthetic_density(mol)
    keys = ("rho", "xc_point", "rho_lm", "poisson", "interp", "contract", "comm", "total")
    for it in range(iters):
        t = time.time()
        J, XC, exc, nel = g.iteration(P)
        w = time.time() - t
        tm = g.timings()
        print("  pair iter %d wall %.2f ms" % (it, w * 1e3), {k: round(tm[k], 3) for k in keys}, flush=True)
    for it in range(iters):
        t = time.time()
        F, ej, exc2, nel2 = g.fock(P)
        w = time.time() - t
        tm = g.timings()
        print("  fock iter %d wall %.2f ms" % (it, w * 1e3), {k: round(tm[k], 3) for k in keys}, flush=True)
    print("  max|F-(2J+XC)| %.2e  ej-2tr(PJ) %.2e  exc %s nel %.9f" % (np.max(np.abs(F - (2 * J + XC))), ej - 2 * np.trace(P @ J), exc == exc2, nel2))
    nb = mol.nbf
    print("  contract: pair %.2f TFLOP/s, fock %.2f TFLOP/s of 37.05" % (0, g.npoints * nb * (nb + 1) / (tm["contract"] * 1e-3) / 1e12))
    # device SCF step with a synthetic H / X (orthonormal X => F' = X^T F X): algebra timing only
    rng = np.random.default_rng(1)
    A = rng.standard_normal((nb, nb))
    H = -(A @ A.T) / nb - np.diag(np.linspace(0, 20, nb)[::-1])
    Q, _ = np.linalg.qr(rng.standard_normal((nb, nb)))
    g.scf_init(H, Q, mol.nelec // 2, 0.5)
    r = g.scf_step(include_xc=False)
    for it in range(3):
        t = time.time()
        r = g.scf_step()
        print("  scf step wall %.2f ms" % ((time.time() - t) * 1e3), r, flush=True)
    g.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "h2o8", int(sys.argv[2]) if len(sys.argv) > 2 else 1)
