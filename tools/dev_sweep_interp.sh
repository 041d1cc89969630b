export DFTGRID_DEVELOPER=1  # the library honours its developer switches only with this set
for R in 2 3 4; do DFTGRID_INTERP_R=$R python tools/dev_prof_fock.py h2o64 3 2>&1 | tail -n 1 | grep -o "'interp': [0-9.]*" | sed "s/^/R=$R /"; done
DFTGRID_INTERP_MINB5=1 python tools/dev_prof_fock.py h2o64 3 2>&1 | tail -n 1 | grep -o "'interp': [0-9.]*" | sed "s/^/R=2 MINB5 /"
for R in 2 3 4; do DFTGRID_INTERP_R=$R python tools/dev_prof_fock.py c40h82 2 2>&1 | tail -n 1 | grep -o "'interp': [0-9.]*" | sed "s/^/c40h82 1e7 R=$R /"; done
