"""Race evidence for the two warp-specialised tensor kernels (k_rho_tma, k_contract_tma).  compute-sanitizer's racecheck
does not model mbarrier-ordered cp.async.bulk traffic (it reports the producer's writes against the DMMA warps' fragment
loads), so two independent lines of evidence are kept in the suite instead:
 1. timing perturbation: pseudo-random delays injected into the producer and / or consumer warps (dftgrid_debug_set_stress,
    compiled only into the test build libdftgrid_stress.so) must not change a single bit of rho, J, XC, F — an ordering bug would;
 2. compute-sanitizer --tool memcheck on a whole small iteration (out-of-bounds / misaligned bulk copies, bad peer pointers)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, grid_params, load_golden, system_from_golden

pytestmark = pytest.mark.gpu


STRESS_LIB = os.path.join(ROOT, "dftcxx_b200", "libdftgrid_stress.so")

WORKER = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np
from common import grid_params, load_golden, system_from_golden
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import WORKLOADS, synthetic_density

def check(mg, P, modes, reps):
    J0, XC0, exc0, nel0 = mg.iteration(P)
    rho0 = mg.get_densities()
    F0, ej0, _, _ = mg.fock(P)
    for mode in modes:
        mg.debug_set_stress(mode)
        for _ in range(reps):
            J, XC, exc, nel = mg.iteration(P)
            assert np.array_equal(J, J0) and np.array_equal(XC, XC0) and exc == exc0 and nel == nel0, mode
            assert np.array_equal(mg.get_densities(), rho0), mode
            F, ej, _, _ = mg.fock(P)
            assert np.array_equal(F, F0) and ej == ej0, mode
    mg.debug_set_stress(0)

name = sys.argv[1]
if name in WORKLOADS:
    fac, prm = WORKLOADS[name]
    mol = fac()
    mg = MolecularGrid(mol); mg.set_grid_parameters(*prm); mg.create_grid()
    check(mg, synthetic_density(mol), (3,), 1)
else:
    g = load_golden(name)
    mg = MolecularGrid(system_from_golden(g)); mg.set_grid_parameters(*grid_params(g)); mg.create_grid()
    check(mg, g["P"], (1, 2, 3), 2)
mg.close()
print("STRESS_OK", name)
"""


def run_worker(name):
    if not os.path.exists(STRESS_LIB):
        pytest.skip("libdftgrid_stress.so not built (make -C dftcxx_b200/csrc stress)")
    code = WORKER % dict(root=ROOT, tests=os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code, name], capture_output=True, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, DFTGRID_LIB=STRESS_LIB))
    assert r.returncode == 0 and "STRESS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("name", ["benzene_p631_fine", "h2o8_p631_fine"])
def test_pipeline_results_are_independent_of_producer_and_consumer_timing(name):
    run_worker(name)


def test_large_tiles_under_stress():
    """(H2O)32: nb = 416 = several tile pairs, an edge tile, multi-segment stream-K CTAs, three-stage pipelines running for
    thousands of stages per CTA."""
    run_worker("h2o32")


def test_production_library_has_no_stress_hook():
    from dftcxx_b200.grid import GridError, MolecularGrid

    g = load_golden("h2o_sto3g")
    mg = MolecularGrid(system_from_golden(g))
    mg.set_grid_parameters(*grid_params(g))
    mg.create_grid()
    try:
        with pytest.raises(GridError, match="no stress hook"):
            mg.debug_set_stress(3)
    finally:
        mg.close()


def test_memcheck_clean_on_a_small_iteration():
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from common import load_golden, system_from_golden, grid_params\n"
        "from dftcxx_b200.grid import MolecularGrid\n"
        "g = load_golden('h2o_sto3g')\n"
        "mg = MolecularGrid(system_from_golden(g)); mg.set_grid_parameters(*grid_params(g)); mg.create_grid()\n"
        "J, XC, exc, nel = mg.iteration(g['P'])\n"
        "F, ej, _, _ = mg.fock(g['P'])\n"
        "nocc = int(round(float(g['nel']))) // 2\n"
        "mg.scf_init(g['scf_H'], g['scf_X'], nocc, 0.5); mg.scf_step(False); mg.scf_step(True)\n"
        "assert np.max(np.abs(J - g['J'])) < 1e-10 and np.max(np.abs(F - 2 * g['J'] - g['XC'])) < 3e-10\n"
        "S, T, V = mg.one_electron(); assert np.max(np.abs(S - g['scf_S'])) < 1e-13 and np.max(np.abs(T + V - g['scf_H'])) < 1e-11\n"
        "pos, rho, grad = mg.rectangular_density(5.0, 6, g['P']); assert np.isfinite(grad).all() and rho.min() >= 0.0\n"
        "mg.close(); print('SANITIZED_RUN_OK')\n" % (ROOT, os.path.join(ROOT, "tests")))
    r = subprocess.run([exe, "--tool", "memcheck", "--error-exitcode", "9", sys.executable, "-c", code], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "SANITIZED_RUN_OK" in r.stdout and "ERROR SUMMARY: 0 errors" in r.stdout, r.stdout[-3000:]
