#!/bin/bash
# multi-GPU call (final build of round 2): real-rank tests + one bench line per workload at N = $1; every multi-rank
# command under `timeout`
N=${1:-2}
WL=${2:-c3}
STEPS=${3:-20}
TAG=${4:-r3m}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s --timeout 300 -k "real or ranks_nccl" > gpurun_out/${TAG}_pytest_$N.log 2>&1
echo "pytest rc=$?"
grep -a "passed\|failed" gpurun_out/${TAG}_pytest_$N.log | tail -3
for w in $WL; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --workload $w --steps $STEPS --warmup 5 2> gpurun_out/${TAG}_bench_${w}_${N}gpu.err | grep -a '^{' | tail -n 1 > gpurun_out/${TAG}_bench_${w}_${N}gpu.json
  echo "bench $w x$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${w}_${N}gpu.json").read().strip().splitlines()[-1])
    print("  ms", round(d["ms_per_step"],4), "flushed", (d.get("flushed_single_replays") or {}).get("ms_per_step"), "sust", round(d["sustained"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"])
    print("  kernels", d["phase_us"].get("kernel_begin_end_us_rank0"))
    print("  phases", d["phase_us"].get("end_of_phase_us_since_step_start_max_over_ranks"))
except Exception as e:
    print("  no line:", e)
PY
  grep -a "Error\|error" gpurun_out/${TAG}_bench_${w}_${N}gpu.err | head -3
done
