#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src).  Run in the build container; the fixtures are committed so the
parity tests need neither /root/reference nor oracle/_ref at run time.

Per molecule:  the flat system the reference parsed, grid parameters, a synthetic symmetric P (seed 20240607),
and the reference's outputs for that P — J, XC, E_xc, electron count before/after the rescale, rho_lm, U_lm and
per-atom charges in full; positions, weights, Becke weights, rho, V, V_fuzzy and Phi on a strided sample of
points (`idx`).  For the small molecules additionally the reference's own SCF: total energy and its components at
every iteration, and (P, J, XC) at iteration `scf_probe_iter` for a realistic density matrix.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dftcxx_b200.molecule import DATA  # noqa: E402
from dftcxx_b200.systems import synthetic_density  # noqa: E402
from oracle.refpy import Ref  # noqa: E402

#            name                   stride  scf iterations (0 = grid only)
CASES = [("h2o_sto3g", 1, 14), ("h2o_p631", 3, 17), ("he_sto3g", 1, 6), ("co_sto3g_coarse", 1, 8),
         ("h2_sto3g_ultrafine", 7, 6), ("ch4_p631_fine", 11, 16), ("ethane_p631_fine", 17, 6),
         ("benzene_p631_fine", 37, 4), ("ch4_p631_dense422", 997, 0), ("h2o8_p631_fine", 211, 5)]


class _M:
    def __init__(self, s, nelec):
        self.nbf = len(s["bf_nprim"])
        self.nelec = nelec


def make(name, stride, nscf):
    path = os.path.join(DATA, "molecules", name + ".in")
    r = Ref(path, full=False)
    s = r.system()
    out = {k: np.asarray(v) for k, v in s.items()}
    P = synthetic_density(_M(s, r.nelec))
    out["P"] = P
    idx = np.arange(0, r.npts, stride)
    xyz, w, wb = r.grid()
    out.update(idx=idx, pts=xyz[idx], w=w[idx], wb=wb[idx], phi=r.amplitudes()[idx])
    out["wsum"] = np.array([w.sum(), wb.sum()])
    out["nel_raw"] = np.array(r.set_density_raw(P))
    out["rho_raw"] = r.densities()[idx]
    r.set_density(P)
    out["rho"] = r.densities()[idx]
    out["nel"] = np.array(r.electron_count())
    out["J"] = r.hartree()
    hi = r.hartree_intermediates()
    rad_idx = np.arange(0, r.nrad, 8 if r.nrad > 100 else 1)  # thin the radial tables of the 422-node case
    out.update(rad_idx=rad_idx, rho_lm=hi["rho_lm"][:, rad_idx], U_lm=hi["U_lm"][:, rad_idx], q=hi["q"], V=hi["V"][idx],
               V_fuzzy=hi["V_fuzzy"][idx])
    XC, exc = r.xc()
    out.update(XC=XC, exc=np.array(exc))
    r.close()
    if nscf:
        f = Ref(path, full=True)
        e0 = f.energies()
        out["scf_H"] = f.matrix("H")
        out["scf_S"] = f.matrix("S")
        out["scf_X"] = f.matrix("X")
        out["scf_P0"] = f.matrix("P")
        out["scf_J0"] = f.matrix("J")
        out["scf_enuc"] = np.array(e0["enuc"])
        rows = []
        probe = min(2, nscf)
        for it in range(1, nscf + 1):
            f.scf_step()
            e = f.energies()
            rows.append([e["et"], e["exc"], e["e_one"], e["e_j"], e["nel"]])
            if it == probe:
                out["scf_probe_iter"] = np.array(it)
                out["scf_P"] = f.matrix("P")
                out["scf_J"] = f.matrix("J")
                out["scf_XC"] = f.matrix("XC")
        out["scf_energies"] = np.array(rows)
        f.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "points", len(idx), "of", len(w), "->", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB",
          ("E_final %.7f" % out["scf_energies"][-1, 0]) if nscf else "")


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, stride, nscf in CASES:
        if not only or name in only:
            make(name, stride, nscf)
