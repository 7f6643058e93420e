#!/bin/bash
# GPU test suite + quick product bench lines (profile mode: no CPU baseline, no sustained region)
mkdir -p gpurun_out
TAG=${1:-r3i}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for wl in ${WLS:-c3 c4 c5 c2 c1}; do
  timeout 400 python bench.py --workload $wl --profile-mode --steps 30 --warmup 5 2> gpurun_out/${TAG}_bench_$wl.err | tail -n 1 > gpurun_out/${TAG}_bench_$wl.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_$wl.json"))
print("$wl", round(d["ms_per_step"], 4), "ms; frac", d.get("roofline", {}).get("frac_of_T_roof", d.get("frac_of_roofline")), json.dumps(d.get("kernel_begin_end_us")))
PY
done
