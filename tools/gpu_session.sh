#!/bin/bash
# One parametrised GPU-box session (replaces the per-call lab-notebook scripts of round 1).
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh golden tests tests_sub perf'
# Every step logs to gpurun_out/<step>.log and never aborts the following steps.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/sparse-videogen_b200:$PYTHONPATH
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt 2>&1
for step in "$@"; do
  t0=$(date +%s)
  case "$step" in
    golden)     timeout 900 python tests/golden/make_golden_gpu.py > gpurun_out/golden.log 2>&1 ;;
    tests)      timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1 ;;
    tests_x)    timeout 1500 python -m pytest tests -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1 ;;
    tests_new)  timeout 1200 python -m pytest tests/test_reference_golden_gpu.py tests/test_configs_gpu.py tests/test_dropin_gpu.py -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/tests_new.log 2>&1 ;;
    tests_sub)  SVGB_ATTN_SUB=1 timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py tests/test_configs_gpu.py -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/tests_sub.log 2>&1 ;;
    sanitizer)  for tool in racecheck synccheck; do
                  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_attention_gpu.py -q -m gpu -p no:cacheprovider \
                    -k "selftest or (band and bf16) or empty_rows or shd_layout" > gpurun_out/sanitizer_$tool.log 2>&1
                done ;;
    perf)       PERF_TAG=${PERF_TAG:-r02} timeout 600 python tools/attn_perf.py > gpurun_out/perf.log 2>&1
                PERF_TAG=${PERF_TAG:-r02} timeout 600 python tools/ab_varblock.py >> gpurun_out/perf.log 2>&1 ;;
    perf_sub)   SVGB_ATTN_SUB=1 PERF_TAG=sub PERF_BAND_ONLY=1 timeout 300 python tools/attn_perf.py > gpurun_out/perf_sub.log 2>&1 ;;
    stages)     timeout 900 python tools/stage_probe.py > gpurun_out/stages.log 2>&1 ;;
    bench)      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ;;
    bench_ref)  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ;;
    trace)      L=$PWD/sparse-videogen_b200/svgb200/_lib
                for c in vb band; do for i in 0 1; do
                  [ -f $L/libsvgb200_trace$i.so ] && SVGB200_LIB=$L/libsvgb200_trace$i.so TRACE_CASE=$c TRACE_TAG=item$i timeout 200 python tools/attn_trace.py >> gpurun_out/trace.log 2>&1
                done; done ;;
    smoke)      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ;;
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_bench.log 2>&1 ;;
    ncu_band)   PROFILE_MODE=band PROFILE_H=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/attn_band python tools/profile_attn.py > gpurun_out/ncu_band.log 2>&1 ;;
    ncu_vb)     PROFILE_MODE=varblock PROFILE_QCKC=400,1000 PROFILE_H=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/attn_varblock python tools/profile_attn.py > gpurun_out/ncu_vb.log 2>&1 ;;
    *)          if [ -f "$step" ]; then timeout 1200 bash "$step" > "gpurun_out/$(basename "$step").log" 2>&1; else echo "unknown step $step"; fi ;;
  esac
  echo "$step rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/session.log
done
tail -n 3 gpurun_out/*.log 2>/dev/null | tail -n 60
