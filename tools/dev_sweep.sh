#!/bin/bash
# developer sweep: env-var variants of the contraction schedule
export DFTGRID_DEVELOPER=1  # the library honours its developer switches only with this set
for dc in 15 16 17; do for nc in 11 12 13; do echo -n "DIAG_COST=$dc NARROW_COST=$nc "; DFTGRID_DIAG_COST=$dc DFTGRID_NARROW_COST=$nc timeout 200 python tools/dev_perf.py h2o64 2>&1 | grep "iter 2" | sed "s/.*'contract': \([0-9.]*\).*/contract \1/"; done; done
