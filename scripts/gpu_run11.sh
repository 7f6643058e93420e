#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
for w in c3 c5; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_$w.json 2> gpurun_out/r2k_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2k_bench_$w.json").read().strip().splitlines()[-1])
    print("$w ms", round(d["ms_per_step"],4), "sust", round(d["sustained"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,4) for k,v in (d.get("stage_ms") or {}).items() if k in ("spring_mid","knn_bound","knn_scan","knn_select","update")}); print("  ", d.get("kernel_begin_end_us")); print("  ", d["details"]["graph_generation_s"], d["details"]["constructor_s"], d["roofline_other"]["iteration"])
except Exception as e:
    print("$w failed", e)
PY
  tail -3 gpurun_out/r2k_bench_$w.err | cut -c1-300
done
