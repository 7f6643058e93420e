#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log
python bench.py --workload c2 --sample-size 1000000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_c2_fullknn.json 2> gpurun_out/r2l_bench_c2_fullknn.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_bench_c2_fullknn.json").read().strip().splitlines()[-1])
    print("c2 full-KNN ms", round(d["ms_per_step"],3), "roof", d["roofline"]["frac"], d["roofline_other"].get("iteration"))
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/r2l_bench_c2_fullknn.err | cut -c1-300
