#!/usr/bin/env python3
"""Golden fixtures at BASELINE.json's LARGE configurations, from the UNMODIFIED reference (oracle/_ref parity build).
Run in the build container (CPU only; minutes to hours — start it in the background):

    python tests/golden/make_golden_large.py h2o32_p631_fine c40h82_p631_fine h2o64_p631_fine

Same content as make_golden.py, thinned so the files stay committable: symmetric matrices are stored as packed
upper triangles (`*_triu`, expanded again by tests/common.load_golden), the synthetic P is regenerated from its seed
by the tests (a checksum is stored), Phi is kept on a coarser point sample (`idx_phi`) than rho / V / weights (`idx`).
One reference object serves both parts (its serial grid construction takes up to ~20 min at (H2O)64): first the
reference's own SCF iterations (total energy and components per iteration, (P, J, XC) at iteration `scf_probe_iter`),
then the fixed-P pass.  Wall-clock of every phase is stored under `ref_seconds_*` (parity build, this container's
cores — informative only)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dftcxx_b200.molecule import DATA  # noqa: E402
from dftcxx_b200.systems import synthetic_density  # noqa: E402
from oracle.refpy import Ref  # noqa: E402

#            name                stride  phi stride  scf iterations
CASES = [("h2o32_p631_fine", 149, 1493, 3), ("c40h82_p631_fine", 173, 1733, 2), ("h2o64_p631_fine", 293, 2931, 0)]


class _M:
    def __init__(self, nbf, nelec):
        self.nbf = nbf
        self.nelec = nelec


def triu(M):
    return np.ascontiguousarray(M[np.triu_indices(M.shape[0])])


def make(name, stride, phi_stride, nscf, outdir=HERE):
    path = os.path.join(DATA, "molecules", name + ".in")
    out = {}
    t0 = time.time()
    r = Ref(path, full=nscf > 0)
    out["ref_seconds_open"] = np.array(time.time() - t0)
    print(name, "opened in %.0f s: %d atoms, %d bf, %d points" % (time.time() - t0, r.natoms, r.nbf, r.npts), flush=True)
    s = r.system()
    out.update({k: np.asarray(v) for k, v in s.items()})
    if nscf:
        e0 = r.energies()
        out["scf_H_triu"] = triu(r.matrix("H"))
        out["scf_enuc"] = np.array(e0["enuc"])
        rows, secs = [], []
        probe = min(2, nscf)
        for it in range(1, nscf + 1):
            t = time.time()
            r.scf_step()
            secs.append(time.time() - t)
            e = r.energies()
            rows.append([e["et"], e["exc"], e["e_one"], e["e_j"], e["nel"]])
            print(name, "scf iteration", it, "E = %.9f  (%.0f s)" % (e["et"], secs[-1]), flush=True)
            if it == probe:
                out["scf_probe_iter"] = np.array(it)
                out["scf_P_triu"] = triu(r.matrix("P"))
                out["scf_J_triu"] = triu(r.matrix("J"))
                out["scf_XC_triu"] = triu(r.matrix("XC"))
        out["scf_energies"] = np.array(rows)
        out["ref_seconds_scf"] = np.array(secs)
    P = synthetic_density(_M(r.nbf, r.nelec))
    out["P_checksum"] = np.array([P.sum(), np.abs(P).sum(), np.trace(P)])
    idx = np.arange(0, r.npts, stride)
    idx_phi = np.arange(0, r.npts, phi_stride)
    xyz, w, wb = r.grid()
    out.update(idx=idx, idx_phi=idx_phi, pts=xyz[idx], w=w[idx], wb=wb[idx])
    phi = r.amplitudes()
    out["phi"] = phi[idx_phi]
    del phi
    out["wsum"] = np.array([w.sum(), wb.sum()])
    t = time.time()
    out["nel_raw"] = np.array(r.set_density_raw(P))
    out["rho_raw"] = r.densities()[idx]
    r.set_density(P)
    out["ref_seconds_density"] = np.array(time.time() - t)
    out["rho"] = r.densities()[idx]
    out["nel"] = np.array(r.electron_count())
    t = time.time()
    out["J_triu"] = triu(r.hartree())
    out["ref_seconds_hartree"] = np.array(time.time() - t)
    print(name, "hartree %.0f s" % (time.time() - t), flush=True)
    hi = r.hartree_intermediates()
    out.update(rad_idx=np.arange(r.nrad), rho_lm=hi["rho_lm"], U_lm=hi["U_lm"], q=hi["q"], V=hi["V"][idx], V_fuzzy=hi["V_fuzzy"][idx])
    t = time.time()
    XC, exc = r.xc()
    out["ref_seconds_xc"] = np.array(time.time() - t)
    out.update(XC_triu=triu(XC), exc=np.array(exc))
    r.close()
    f = os.path.join(outdir, name + ".npz")
    np.savez_compressed(f, **out)
    print(name, "->", os.path.getsize(f) // 1024, "KiB, total %.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    outdir = HERE
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            outdir = a[6:]
    for name, stride, phi_stride, nscf in CASES:
        if not only or name in only:
            make(name, stride, phi_stride, nscf, outdir)
