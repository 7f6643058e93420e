#!/bin/bash
# final single-GPU lines of round 2 + ncu evidence
cd /root/repo; mkdir -p gpurun_out
python bench.py --workload c3 --steps 30 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; echo "c3 rc=$?"
python bench.py --workload c3 --sampler torch --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_1gpu_torch_sampler.json 2> gpurun_out/r02_bench_c3_1gpu_torch_sampler.err; echo "c3 torch rc=$?"
for w in c1 c2; do python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/r02_bench_${w}_1gpu.err; echo "$w rc=$?"; done
for w in c4 c5; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/r02_bench_${w}_1gpu.err; echo "$w rc=$?"; done
python bench.py --impl reference --workload c3 --steps 3 --warmup 1 > gpurun_out/r02_bench_c3_reference_arm_cpu.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r02_plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r02_plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'knn_prep|knn_scan|knn_select|spring_csr|update_pass' -s 36 -c 12 \
    -o gpurun_out/r02_prof_c3 python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
