#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
for w in c3 c4; do
python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_$w.json 2> gpurun_out/r2n_bench_$w.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench_$w.json").read().strip().splitlines()[-1])
print("$w ms", round(d["ms_per_step"],4), "sust", round(d["sustained"]["ms_per_step"],4), "iter", d["roofline_other"]["iteration"]); print("  ", d.get("kernel_begin_end_us"))
PY
done
