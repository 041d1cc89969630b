#!/bin/bash
# Round-2 evidence run (one GPU, under gpurun): bench lines, ncu launch list, ncu --set full of the iteration kernels at
# (H2O)64, C40H82/fine, C40H82 at 10^7 points (whole and a 1/8 shard) and of the SCF algebra kernels.  Outputs: gpurun_out/r02_*.
set -x
O=gpurun_out
python bench.py --steps 10 > $O/r02_bench_h2o64.json 2> $O/r02_bench_h2o64.err
python bench.py --workload c40h82_fine --steps 10 --no-cpu-baseline > $O/r02_bench_c40h82_fine.json 2> /dev/null
python bench.py --workload h2o32 --steps 10 --no-cpu-baseline > $O/r02_bench_h2o32.json 2> /dev/null
python bench.py --workload benzene --steps 20 > $O/r02_bench_benzene.json 2> /dev/null
python bench.py --workload c40h82 --steps 5 --no-cpu-baseline --no-scf > $O/r02_bench_c40h82_1e7pts.json 2> $O/r02_bench_c40h82_1e7pts.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_h2o64.csv python tools/dev_prof_fock.py h2o64 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_contract_tma|k_rho_tma|k_interp_bin" -s 9 -c 3 -o $O/r02_prof_h2o64 python tools/dev_prof_fock.py h2o64 2 > $O/r02_prof_h2o64.log 2>&1
ncu --set full --clock-control none -k regex:"k_contract_tma|k_rho_tma|k_interp_bin" -s 9 -c 3 -o $O/r02_prof_c40fine python tools/dev_prof_fock.py c40h82_fine 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_contract_tma|k_rho_tma|k_interp_bin" -s 9 -c 3 -o $O/r02_prof_c40_1e7 python tools/dev_prof_fock.py c40h82 2 > $O/r02_prof_c40_1e7.log 2>&1
FAKE_RANK=0 FAKE_NRANKS=8 ncu --set full --clock-control none -k regex:"k_contract_tma|k_rho_tma|k_interp_bin" -s 9 -c 3 -o $O/r02_prof_c40_1e7_shard0of8 python tools/dev_prof_fock.py c40h82 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_gemm_nn" -s 4 -c 2 -o $O/r02_prof_scf python tools/dev_prof_scf.py h2o64 > $O/r02_prof_scf.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_scf.csv python tools/dev_prof_scf.py h2o64 > /dev/null 2>&1
