"""Developer sweep of the contraction cost-model constants (environment switches) on a workload: prints the contraction time."""
import itertools
import os

os.environ.setdefault("DFTGRID_DEVELOPER", "1")  # developer script: the library's A/B switches are live
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
grid = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    grid[k] = v.split(",")
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    env = dict(os.environ)
    for k, v in zip(keys, combo):
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dev_prof_fock.py"), name, "4"], capture_output=True, text=True, env=env)
    m = re.search(r"'contract': ([0-9.]+)", r.stdout)
    print(dict(zip(keys, combo)), "contract ms", m.group(1) if m else r.stderr[-300:], flush=True)
