#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
for w in c3 c2 c1; do
  python bench.py --workload $w --steps 20 --warmup 5 --profile-mode > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err
done
python scripts/rank_share_profile.py c3 8 10 > gpurun_out/r2f_rankshare_c3_w8.log 2>&1; grep -v Warning gpurun_out/r2f_rankshare_c3_w8.log | tail -2 | cut -c1-600
