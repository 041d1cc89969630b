"""Developer script: per-phase device timings of build + iteration for a synthetic workload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import WORKLOADS, synthetic_density

def main(name, iters=3):
    fac, (nr, lo, lm) = WORKLOADS[name]
    mol = fac()
    rank, nranks = int(os.environ.get("FAKE_RANK", "0")), int(os.environ.get("FAKE_NRANKS", "1"))
    g = MolecularGrid(mol, rank=rank, nranks=nranks); g.set_grid_parameters(nr, lo, lm)
    if nranks > 1:
        os.environ["DFTGRID_DEBUG_SKIP_COMM"] = "1"  # time one shard's kernels on a single GPU (results are partial sums)
    t = time.time(); g.create_grid(comm_id=False if nranks > 1 else None)
    t1 = time.time() - t
    tm = g.timings()
    print(name, "natoms", mol.natoms, "nbf", mol.nbf, "npts", g.npoints, "create_grid wall %.3fs" % t1,
          {k: round(tm[k], 3) for k in ("points", "becke", "phi")})
    print("  phi evals/s %.3e  becke pairs/s %.3e" % (g.npoints * mol.nbf / (tm["phi"] * 1e-3), g.npoints * mol.natoms ** 2 / (tm["becke"] * 1e-3)))
    P = synthetic_density(mol)
    for it in range(iters):
        t = time.time(); J, XC, exc, nel = g.iteration(P); t1 = time.time() - t
        tm = g.timings()
        print("  iter %d wall %.2f ms" % (it, t1 * 1e3), {k: round(v, 3) for k, v in tm.items() if k not in ("points", "becke", "phi")}, "nel %.6f exc %.6f trJ %.9f" % (nel, exc, float(np.trace(J))))
    F = 2.0 * g.npoints * g.nbf ** 2
    print("  rho: %.2f TFLOP/s (full 2*N*nb^2)   contract: %.2f TFLOP/s (2 sym matrices, N*nb*(nb+1) each)" % (
        F / (tm["rho"] * 1e-3) / 1e12, 2.0 * g.npoints * g.nbf * (g.nbf + 1) / (tm["contract"] * 1e-3) / 1e12))

if __name__ == "__main__":
    for n in sys.argv[1:] or ["h2o8"]:
        main(n)
