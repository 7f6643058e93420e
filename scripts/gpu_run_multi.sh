#!/bin/bash
# multi-GPU call: real-rank tests + bench lines at N = $1 (default 2); every multi-rank command under `timeout`
N=${1:-2}
WL=${2:-c3}
STEPS=${3:-20}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_gpus_$N.txt 2>&1
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s --timeout 300 -k "real or ranks_nccl" > gpurun_out/r2m_pytest_$N.log 2>&1
echo "pytest rc=$?"
grep -a "flags\|passed\|failed\|exchange" gpurun_out/r2m_pytest_$N.log | tail -20
for w in $WL; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $w --steps $STEPS --warmup 5 > gpurun_out/r2m_bench_${w}_${N}gpu.json 2> gpurun_out/r2m_bench_${w}_${N}gpu.err
  echo "bench $w x$N rc=$?"
  tail -c 1500 gpurun_out/r2m_bench_${w}_${N}gpu.json
  grep -a "Error\|error" gpurun_out/r2m_bench_${w}_${N}gpu.err | head -5
done
