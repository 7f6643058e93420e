#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
for R in 2 8; do
  python scripts/rank_share_profile.py c3 $R 10 > gpurun_out/r2d_rankshare_c3_w$R.log 2>&1
  echo "rankshare $R rc=$?"; tail -2 gpurun_out/r2d_rankshare_c3_w$R.log
done
python scripts/rank_share_profile.py c3 8 2 > gpurun_out/r2d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn_|spring|update_|topk|rows_' -c 60 --csv --log-file gpurun_out/r2d_launches_rankshare_w8.csv \
    python scripts/rank_share_profile.py c3 8 2 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"
for w in c3 c2; do
  python bench.py --workload $w --steps 20 --warmup 5 --profile-mode > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err
done
