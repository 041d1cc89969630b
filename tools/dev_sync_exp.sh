#!/bin/bash
# per-launch DRAM / time of the LAST 3 contraction launches (the iterations) for a few soft-lockstep settings
export DFTGRID_DEVELOPER=1  # the library honours its developer switches only with this set
for cfg in "0 2" "24 2" "12 3" "48 2" "24 4" "12 2"; do set -- $cfg
  echo "== SYNC_MB=$1 LEAD=$2"
  DFTGRID_CON_SYNC_MB=$1 DFTGRID_CON_SYNC_LEAD=$2 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_contract_tma --clock-control none -c 14 python tools/dev_prof_fock.py h2o64 3 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate" | tail -6 | tr -s ' ' | tr '\n' ' '; echo
  echo -n "   eager: "; DFTGRID_CON_SYNC_MB=$1 DFTGRID_CON_SYNC_LEAD=$2 python tools/dev_prof_fock.py h2o64 6 2>&1 | grep "contract" | sed "s/.*'contract': \([0-9.]*\).*/contract \1 ms/"
done
