#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
echo "== tail kernel trace (item 0 of the tail list)"
SVGB200_LIB=$L/libsvgb200_trace0.so TRACE_CASE=vb TRACE_TAG=tail timeout 200 python tools/attn_trace.py 2>&1 | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/attn_trace_vb_tail.json'))
for r in d['rows'][10:16]: print(r['j'], r['w4'])
PY
echo "== dynmap"
timeout 120 python tools/profile_dynmap.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dynmap -c 1 -o gpurun_out/dynmap python tools/profile_dynmap.py > gpurun_out/ncu_dynmap.log 2>&1
python tools/ncu_summary.py gpurun_out/dynmap.ncu-rep gpurun_out/dynmap_summary.json 2>&1 | tail -45
