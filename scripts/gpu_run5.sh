#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 python scripts/rank_share_profile.py tiny 2 2 > gpurun_out/r2e_rs_tiny.log 2>&1; echo "tiny rc=$?"; grep -v Warning gpurun_out/r2e_rs_tiny.log | tail -3 | cut -c1-400
for R in 2 8; do
  python scripts/rank_share_profile.py c3 $R 10 > gpurun_out/r2e_rankshare_c3_w$R.log 2>&1
  echo "rankshare $R rc=$?"; grep -v Warning gpurun_out/r2e_rankshare_c3_w$R.log | tail -2 | cut -c1-600
done
python scripts/rank_share_profile.py c3 8 2 > gpurun_out/r2e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn_|spring|update_|topk|rows_' -c 60 --csv --log-file gpurun_out/r2e_launches_rankshare_w8.csv \
    python scripts/rank_share_profile.py c3 8 2 > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"
