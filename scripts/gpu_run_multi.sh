#!/bin/bash
# multi-GPU call: real-rank tests + bench lines at N = $1 (default 2); every multi-rank command under `timeout`
N=${1:-2}
WL=${2:-c3}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_gpus_$N.txt 2>&1
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q --timeout 600 -k "real or ranks_nccl" > gpurun_out/r2m_pytest_$N.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/r2m_pytest_$N.log
for w in $WL; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $w --steps 20 --warmup 5 > gpurun_out/r2m_bench_${w}_${N}gpu.json 2> gpurun_out/r2m_bench_${w}_${N}gpu.err
  echo "bench $w x$N rc=$?"
  tail -c 600 gpurun_out/r2m_bench_${w}_${N}gpu.json
done
