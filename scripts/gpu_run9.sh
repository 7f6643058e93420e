#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python scripts/scan_cta_dump.py c3 > gpurun_out/r2i_scan_ctas_c3.log 2>&1; echo rc=$?; cat gpurun_out/r2i_scan_ctas_c3.log | cut -c1-1500
