#!/usr/bin/env python3
"""Golden fixture for the density dump: density and density gradient on the reference's
RectangularGrid (src/rectangulargrid.cpp:34-95), from the UNMODIFIED reference (oracle/_ref).  Run in the build container:

    python tests/golden/make_golden_rect.py          # -> tests/golden/rect_density.npz

Cases: h2o / STO-3G on the 5.0 x 15^3 box of the reference's own (commented-out) DFT::finalize (src/dft.cpp:493-497),
benzene / 6-31G on a 12.0 x 9^3 box and CO / STO-3G on a 6.0 x 11^3 box.  (Cartesian D shells, where the reference's
CGF::get_grad drops the factor l of the monomial derivative, start at Sc; the reference's parser stops at Ar,
src/molecule.cpp:278-287, so no reference run can reach them.)
P is the synthetic symmetric density of dftcxx_b200.systems.synthetic_density plus, for h2o, a NON-symmetric P (both terms
of GridPoint::set_gradient are kept).  The first lines of the reference's own dump file pin the text format."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dftcxx_b200.molecule import DATA  # noqa: E402
from dftcxx_b200.systems import synthetic_density  # noqa: E402
from oracle.refpy import Ref  # noqa: E402

CASES = [("h2o_sto3g", 5.0, 15), ("benzene_p631_fine", 12.0, 9), ("co_sto3g_coarse", 6.0, 11)]


class _M:
    def __init__(self, nbf, nelec):
        self.nbf, self.nelec = nbf, nelec


def main():
    out = {"cases": np.array([c[0] for c in CASES])}
    for name, size, dp in CASES:
        r = Ref(os.path.join(DATA, "molecules", name + ".in"))
        for k, v in r.system().items():
            out["%s.%s" % (name, k)] = np.asarray(v)
        P = synthetic_density(_M(r.nbf, r.nelec))
        pos, rho, grad = r.rect_density(size, dp, P)
        out.update({name + ".P": P, name + ".size": np.array(size), name + ".dp": np.array(dp), name + ".pos": pos, name + ".rho": rho,
                    name + ".grad": grad})
        if name == "h2o_sto3g":
            rng = np.random.default_rng(7)
            Pn = P + 0.1 * rng.standard_normal(P.shape)  # not symmetric
            _, rho_n, grad_n = r.rect_density(size, dp, Pn)
            out.update({name + ".P_nonsym": Pn, name + ".rho_nonsym": rho_n, name + ".grad_nonsym": grad_n})
            with tempfile.TemporaryDirectory() as d:
                fn = os.path.join(d, "data.dat")
                r.rect_write(size, dp, P, fn)
                lines = open(fn).read().splitlines()
            assert len(lines) == dp ** 3
            out[name + ".dump_lines"] = np.array(lines[:40] + lines[-40:])
        print(name, "nbf", r.nbf, "points", dp ** 3, "max rho %.4f max |grad| %.4f" % (rho.max(), np.abs(grad).max()))
        r.close()
    f = os.path.join(HERE, "rect_density.npz")
    np.savez_compressed(f, **out)
    print("->", f, os.path.getsize(f) // 1024, "KiB")


if __name__ == "__main__":
    main()
