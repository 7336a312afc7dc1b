#!/bin/bash
cd "$(dirname "$0")/.."
echo "== tests"
timeout 1200 python -m pytest tests/test_attention_gpu.py tests/test_svg2_ops_gpu.py tests/test_reference_golden_gpu.py tests/test_configs_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py -q -m gpu -rf --no-header -p no:cacheprovider 2>&1 | tail -12
timeout 300 python tools/kmeans_probe.py 2>&1 | tail -12
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench16.json 2> gpurun_out/bench16.err; tail -c 600 gpurun_out/bench16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench16.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d.get('svg2_pipeline'))[:3000])
print(json.dumps(d.get('svg2'))[:600])
PY
