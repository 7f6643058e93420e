#!/bin/bash
# ncu evidence of the final round-2 build (C3, 1 GPU): plain run first (must exit 0), then the launch list, then one
# `--set full` capture of two iterations' kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode"
$CMD > gpurun_out/r02f_plain_c3.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_c3.csv $CMD > gpurun_out/r02f_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'knn_prep|knn_scan|knn_select|spring_csr|update_pass' -s 36 -c 12 \
    -o gpurun_out/r02f_prof_c3 $CMD > gpurun_out/r02f_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/r02f_prof_c3.ncu-rep
