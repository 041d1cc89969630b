#!/bin/bash
# Round-2 evidence run (one GPU, under gpurun): bench lines, ncu launch lists, ncu --set full of the iteration kernels at
# (H2O)64, C40H82/fine, C40H82 at 10^7 points (whole and a 1/8 shard) and of the SCF algebra kernels.  The .ncu-rep files are
# summarised on the box (gpurun returns at most 64 MiB) and removed; outputs: gpurun_out/${T}_*.
# usage: tools/run_profiles.sh [tag]   (default tag r02)
export DFTGRID_DEVELOPER=1  # the library honours its developer switches only with this set
set -x
O=gpurun_out
T=${1:-r02}
mkdir -p $O
K='regex:k_contract_tma|k_rho_tma|k_interp_bin'
if [ -z "$PROF_ONLY" ]; then   # PROF_ONLY=1: the ncu captures only
python bench.py --steps 10 > $O/${T}_bench_h2o64.json 2> $O/${T}_bench_h2o64.err
python bench.py --workload c40h82_fine --steps 10 --no-cpu-baseline > $O/${T}_bench_c40h82_fine.json 2> /dev/null
python bench.py --workload h2o32 --steps 10 --no-cpu-baseline > $O/${T}_bench_h2o32.json 2> /dev/null
python bench.py --workload benzene --steps 20 > $O/${T}_bench_benzene.json 2> /dev/null
python bench.py --impl reference --steps 1 --warmup 0 > $O/${T}_bench_reference_h2o64.json 2> /dev/null
python bench.py --workload c40h82 --steps 5 --no-cpu-baseline --no-scf > $O/${T}_bench_c40h82_1e7pts.json 2> $O/${T}_bench_c40h82_1e7pts.err
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${T}_launches_h2o64.csv python tools/dev_prof_fock.py h2o64 3 > /dev/null 2>&1
python tools/ncu_summary.py launches $O/${T}_launches_h2o64.csv $O/${T}_launches_h2o64.txt && rm -f $O/${T}_launches_h2o64.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches_scf.csv python tools/dev_prof_scf.py h2o64 > /dev/null 2>&1
python tools/ncu_summary.py launches $O/${T}_launches_scf.csv $O/${T}_launches_scf_h2o64.txt && rm -f $O/${T}_launches_scf.csv
prof() {  # name, kernel regex, skip, count, command...
  local name=$1 kern=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "$kern" -s $skip -c $cnt -o $O/$name "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py full $O/$name.ncu-rep $O/${name/prof/ncu_full}.txt
  python tools/ncu_hot.py $O/$name.ncu-rep k_ 12 > $O/${name/prof/ncu_hot}.txt 2>&1
  rm -f $O/$name.ncu-rep
}
prof ${T}_prof_h2o64 "$K" 9 3 python tools/dev_prof_fock.py h2o64 2
[ -n "$LIGHT" ] && { ls -la $O; exit 0; }   # LIGHT=1: the (H2O)64 captures only
prof ${T}_prof_c40h82_fine "$K" 9 3 python tools/dev_prof_fock.py c40h82_fine 2
prof ${T}_prof_c40h82_1e7pts "$K" 9 3 python tools/dev_prof_fock.py c40h82 2
FAKE_RANK=0 FAKE_NRANKS=8 prof ${T}_prof_c40h82_1e7pts_shard0of8 "$K" 9 3 python tools/dev_prof_fock.py c40h82 2
prof ${T}_prof_scf_h2o64 "regex:k_gemm_nn|k_gemm_sym32" 4 3 python tools/dev_prof_scf.py h2o64
ls -la $O
