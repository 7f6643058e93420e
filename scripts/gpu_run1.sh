#!/bin/bash
# round 2, GPU call 1 (single GPU): smoke, the GPU test suite, bench lines, launch lists
cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?"
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
echo "bench c3 rc=$?"
python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
python bench.py --workload c3 --sampler torch --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c3_torch.json 2> gpurun_out/r2a_bench_c3_torch.err
echo "benches done"
python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2a_plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2a_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
python bench.py --workload c1 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2a_plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c1.csv \
    python bench.py --workload c1 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2a_ncu_c1.log 2>&1
echo "ncu c1 rc=$?"
