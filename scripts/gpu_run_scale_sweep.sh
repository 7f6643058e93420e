#!/bin/bash
# bound-sample scale sweep (GEM_BOUND_SCALE) on the product library: ms per iteration + kernel stamps
mkdir -p gpurun_out
TAG=${1:-r3g}
for wl in ${WLS:-c3}; do
for sc in ${SCALES:-1 1.5 2 3}; do
  GEM_BOUND_SCALE=$sc timeout 300 python bench.py --workload $wl --profile-mode --steps 30 --warmup 5 2> gpurun_out/${TAG}_bench_${wl}_s$sc.err | tail -n 1 > gpurun_out/${TAG}_bench_${wl}_s$sc.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${wl}_s$sc.json"))
print("$wl scale $sc:", round(d["ms_per_step"], 4), "ms", json.dumps(d.get("kernel_begin_end_us")))
PY
done
done
