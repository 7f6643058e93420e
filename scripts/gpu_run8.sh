#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
for w in c3 c4 c2; do
  python bench.py --workload $w --steps 20 --warmup 5 --profile-mode > gpurun_out/r2h_bench_$w.json 2> gpurun_out/r2h_bench_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_bench_$w.json").read().strip().splitlines()[-1])
print("$w ms", round(d["ms_per_step"],4), {k:round(v,4) for k,v in (d.get("stage_ms") or {}).items() if k in ("spring_mid","knn_bound","knn_scan","knn_select","update")}); print("  ", d.get("kernel_begin_end_us"))
PY
done
