#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s --timeout 250 -k "eight" > gpurun_out/r2t_pytest_$i.log 2>&1
echo "pytest $i rc=$?"
grep -a "flags" gpurun_out/r2t_pytest_$i.log | python -c "
import sys,re
for l in sys.stdin:
    r=l.split(']')[0]
    d=eval(l.split('flags: ')[1])
    print(r, {k:v for k,v in d.items() if 'raw' in k or 'stats' in k or 'replicas' in k or 'pos_rows' in k or 'pos_max' in k or 'replay_pos' in k})
"
done
