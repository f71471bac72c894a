#!/bin/bash
# Round-2 profiling pass on one B200 (run through gpurun; outputs land in gpurun_out/).
# 1. launch list of the benchmark command itself  2. one ncu --set full capture per hot kernel
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-safe-api"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c3.csv $B > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches_bench.err
P="python bench.py --steps 1 --warmup 0 --perms 139 --no-cpu-baseline --no-safe-api --no-parity"
ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 1 -c 1 -f -o gpurun_out/prof_gemm_r2 $P > /dev/null 2> gpurun_out/prof_gemm_r2.err
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_fixup" -s 1 -c 2 -f -o gpurun_out/prof_gather_fixup_r2 $P > /dev/null 2> gpurun_out/prof_gather_r2.err
ncu --set full --clock-control none --import-source on -k regex:k_euclid -c 1 -f -o gpurun_out/prof_stage1_hyper_r2 python tools/kernel_bench.py --only euclid > gpurun_out/prof_stage1_hyper_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hypergeom -c 1 -f -o gpurun_out/prof_hypergeom_r2b python tools/kernel_bench.py --only hypergeom > gpurun_out/prof_hypergeom_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sssp" -s 2 -c 1 -f -o gpurun_out/prof_sssp_r2 python tools/kernel_bench.py --only sssp --configs C3 > gpurun_out/prof_sssp_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
