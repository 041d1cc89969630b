#!/bin/bash
# developer sweep: env-var variants of the contraction schedule and the interpolation kernel
for nc in 5 6 7; do echo "NARROW_COST=$nc"; DFTGRID_NARROW_COST=$nc python tools/dev_perf.py h2o64 2>&1 | grep "iter 2"; done
for v in 0 1 2 3; do echo "INTERP_VARIANT=$v"; DFTGRID_INTERP_VARIANT=$v python tools/dev_perf.py h2o64 2>&1 | grep "iter 2"; done
