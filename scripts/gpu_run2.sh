#!/bin/bash
# round 2, GPU call 2 (single GPU): GPU test suite, bench lines, launch list + one full ncu capture of the hot kernels
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
for w in c3 c2 c1 c4; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err
  echo "bench $w rc=$?"
done
python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2b_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'knn_prep|knn_scan|knn_select|spring_csr|update_pass' -s 40 -c 12 \
    -o gpurun_out/r2b_prof_c3 python bench.py --workload c3 --steps 2 --warmup 3 --profile-mode > gpurun_out/r2b_ncu_c3.log 2>&1
echo "ncu full c3 rc=$?"
