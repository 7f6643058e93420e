#!/bin/bash
# diagnostic builds of the library (-DGEM_SCAN_DIAG ...) -> per-phase %globaltimer points of the scan and
# preparation kernels (scripts/scan_diag.py); the product library is untouched
mkdir -p gpurun_out
TAG=${1:-r2x}
shift
WLS=${@:-c1 c2 ba125000 c3}
for v in diag; do
  timeout 420 python scripts/scan_diag.py graphem_rapids_b200/libgraphem_b200_$v.so $WLS > gpurun_out/${TAG}_scan_$v.log 2>&1
  echo "exit $?" >> gpurun_out/${TAG}_scan_$v.log
done
grep -v "^ *$" gpurun_out/${TAG}_scan_diag.log | tail -n 120
for wl in ${BENCH_WLS:-c3 c2 c1}; do
  timeout 300 python bench.py --workload $wl --profile-mode --steps 30 --warmup 5 2> gpurun_out/${TAG}_bench_$wl.err | tail -n 1 > gpurun_out/${TAG}_bench_$wl.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_$wl.json"))
print("$wl", round(d["ms_per_step"], 4), "ms; parity", d.get("parity"), json.dumps(d.get("kernel_begin_end_us")))
PY
done
