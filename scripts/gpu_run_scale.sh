#!/bin/bash
# one 8-GPU box: the 8-rank test, then bench lines at N = 8 for the given workloads (default c3 c4 c5) and C3 at the given extra N
cd /root/repo
mkdir -p gpurun_out
WLS=${1:-"c3 c4 c5"}
EXTRA_N=${2:-""}
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s --timeout 300 -k "eight" > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?"
grep -a "passed\|failed" gpurun_out/r2s_pytest.log | tail -3
run() {  # workload N
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2952$2 \
      bench.py --gpus $2 --workload $1 --steps 20 --warmup 5 > gpurun_out/r2s_bench_$1_$2gpu.json 2> gpurun_out/r2s_bench_$1_$2gpu.err
  echo "bench $1 x$2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bench_$1_$2gpu.json").read().strip().splitlines()[-1])
    print("  ms", round(d["ms_per_step"],4), "sust", round(d["sustained"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"], d["parity"].get("pos_rel_inf"), "gen_s", d["details"]["graph_generation_s"], "ctor_s", d["details"]["constructor_s"])
    print("  kernels", d["phase_us"].get("kernel_begin_end_us_rank0"))
except Exception as e:
    print("  no line:", e)
PY
  grep -a "Error\|error" gpurun_out/r2s_bench_$1_$2gpu.err | head -3
}
for w in $WLS; do run $w 8; done
for n in $EXTRA_N; do run c3 $n; done
