#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for w in c3 c2; do
  python bench.py --workload $w --steps 20 --warmup 5 --profile-mode > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench_$w.json").read().strip().splitlines()[-1])
print("$w ms", round(d["ms_per_step"],4)); print(d.get("kernel_begin_end_us"))
PY
done
