#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
echo "== base single"; SVGB200_LIB=$L/libsvgb200_trace_base.so TRACE_CASE=single timeout 200 python tools/attn_trace.py | tail -9
echo "== stream single"; SVGB200_LIB=$L/libsvgb200_trace.so TRACE_CASE=single timeout 200 python tools/attn_trace.py | tail -9
