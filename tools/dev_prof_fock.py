"""Developer script: a few fused Fock builds of a workload (for ncu captures of the iteration kernels)."""
import os

os.environ.setdefault("DFTGRID_DEVELOPER", "1")  # developer script: the library's A/B switches are live
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dftcxx_b200.grid import MolecularGrid
from dftcxx_b200.systems import WORKLOADS, synthetic_density

name = sys.argv[1] if len(sys.argv) > 1 else "h2o64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fac, prm = WORKLOADS[name]
mol = fac()
rank, nranks = int(os.environ.get("FAKE_RANK", "0")), int(os.environ.get("FAKE_NRANKS", "1"))
if nranks > 1:
    os.environ["DFTGRID_DEBUG_SKIP_COMM"] = "1"  # time one shard's kernels on a single GPU (results are partial sums)
g = MolecularGrid(mol, rank=rank, nranks=nranks)
g.set_grid_parameters(*prm)
g.create_grid(comm_id=False if nranks > 1 else None)
print("rank", rank, "of", nranks, "local points", g.nloc, "offset", g.point_offset)
P = synthetic_density(mol)
os.environ["DFTGRID_NO_GRAPH"] = "1"
for _ in range(n):
    g.fock(P)
print(g.timings(), "screen fraction", g.screen_fraction())
g.close()
