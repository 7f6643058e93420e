#!/bin/bash
# final single-GPU lines of round 2 (bench lines are never taken under a profiler)
cd /root/repo; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
python bench.py --workload c3 --steps 30 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; echo "c3 rc=$?"
python bench.py --workload c3 --sampler torch --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_1gpu_torch_sampler.json 2> gpurun_out/r02_bench_c3_1gpu_torch_sampler.err; echo "c3 torch rc=$?"
for w in c1 c2; do python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/r02_bench_${w}_1gpu.err; echo "$w rc=$?"; done
for w in c4 c5; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/r02_bench_${w}_1gpu.err; echo "$w rc=$?"; done
python bench.py --workload c2 --sample-size 1000000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_fullknn_1gpu.json 2> gpurun_out/r02_bench_c2_fullknn_1gpu.err; echo "c2 full rc=$?"
python bench.py --impl reference --workload c3 --steps 3 --warmup 1 > gpurun_out/r02_bench_c3_reference_arm_cpu.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
