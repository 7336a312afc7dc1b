#!/bin/bash
cd "$(dirname "$0")/.."
KM_PROBE=assign timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -s 3 -c 1 -o gpurun_out/km_assign python tools/kmeans_probe.py > gpurun_out/ncu_km.log 2>&1
python tools/ncu_summary.py gpurun_out/km_assign.ncu-rep gpurun_out/km_assign_summary.json > gpurun_out/km_assign_summary.txt 2>&1
ncu -i gpurun_out/km_assign.ncu-rep --page source --csv > gpurun_out/km_assign_source.csv 2>/dev/null
ls -la gpurun_out/km_assign*
rm -f gpurun_out/km_assign.ncu-rep
cat gpurun_out/km_assign_summary.txt
