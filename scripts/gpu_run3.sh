#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?"
tail -4 gpurun_out/r2c_pytest.log
for w in c3 c2 c1 c4; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err
  echo "bench $w rc=$?"
done
for sc in 0.5 1 4; do
  GEM_BOUND_SCALE=$sc python bench.py --workload c3 --steps 20 --warmup 5 --profile-mode > gpurun_out/r2c_bench_c3_scale$sc.json 2> gpurun_out/r2c_bench_c3_scale$sc.err
done
python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2c_plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2c_ncu_c3.log 2>&1
echo "ncu rc=$?"
