#!/bin/bash
# one 8-GPU box: real-rank tests at 2/4/8, then the C3 bench at N = 8, 4, 2 (and optional extra workloads at 8)
cd /root/repo
mkdir -p gpurun_out
EXTRA=${1:-}
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s --timeout 300 -k "real or ranks_nccl" > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?"
grep -a "passed\|failed\|exchange" gpurun_out/r2s_pytest.log | tail -8
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --workload c3 --steps 20 --warmup 5 > gpurun_out/r2s_bench_c3_${N}gpu.json 2> gpurun_out/r2s_bench_c3_${N}gpu.err
  echo "bench c3 x$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bench_c3_${N}gpu.json").read().strip().splitlines()[-1])
    print("  ms", round(d["ms_per_step"],4), "sust", round(d["sustained"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"], d["parity"].get("pos_rel_inf"))
    print("  phases", d["phase_us"]["end_of_phase_us_since_step_start_max_over_ranks"])
except Exception as e:
    print("  no line:", e)
PY
done
for w in $EXTRA; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus 8 --workload $w --steps 20 --warmup 5 > gpurun_out/r2s_bench_${w}_8gpu.json 2> gpurun_out/r2s_bench_${w}_8gpu.err
  echo "bench $w x8 rc=$?"; tail -c 400 gpurun_out/r2s_bench_${w}_8gpu.json
done
