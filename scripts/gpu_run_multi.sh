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
grep -a "passed\|failed" gpurun_out/r2m_pytest_$N.log | tail -3
for w in $WL; do
 for MC in 1 0; do
  GEM_MULTICAST=$MC timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MC \
      bench.py --gpus $N --workload $w --steps $STEPS --warmup 5 > gpurun_out/r2m_bench_${w}_${N}gpu_mc$MC.json 2> gpurun_out/r2m_bench_${w}_${N}gpu_mc$MC.err
  echo "bench $w x$N multicast=$MC rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2m_bench_${w}_${N}gpu_mc$MC.json").read().strip().splitlines()[-1])
    print("  ms", round(d["ms_per_step"],4), "sust", round(d["sustained"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"])
    print("  ", d["details"]["parallelism"][:120])
    print("  kernels", d["phase_us"].get("kernel_begin_end_us_rank0"))
except Exception as e:
    print("  no line:", e)
PY
  grep -a "Error\|error" gpurun_out/r2m_bench_${w}_${N}gpu_mc$MC.err | head -3
 done
done
