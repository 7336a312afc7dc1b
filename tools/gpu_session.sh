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
    sanitizer)  for tool in racecheck synccheck memcheck; do
                  timeout 420 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_attention_gpu.py -q -m gpu -p no:cacheprovider --no-header \
                    -k "selftest_tile or transposed_tail or empty_rows or (band_attention and bfloat16)" > gpurun_out/sanitizer_$tool.log 2>&1
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
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_bench.log 2>&1 ;;
    ncu_attn)   # ncu --set full of the band / variable-block / transposed-tail kernels -> summaries only (the three
                # .ncu-rep files, ~40 MB each, exceed gpurun's 64 MiB return limit and are deleted on the box)
                for spec in "band band x attn_fwd_kernel" "varblock varblock 400,1000 attn_fwd_kernel" "tail varblock 400,1000 attn_tail_kernel"; do
                  set -- $spec
                  PROFILE_MODE=$2 PROFILE_QCKC=$3 PROFILE_H=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 \
                    -o gpurun_out/attn_$1 python tools/profile_attn.py > gpurun_out/ncu_$1.log 2>&1
                  python tools/ncu_summary.py gpurun_out/attn_$1.ncu-rep gpurun_out/attn_$1_summary.json > gpurun_out/attn_$1_summary.txt 2>&1
                  rm -f gpurun_out/attn_$1.ncu-rep
                done ;;
    ncu_km)     KM_PROBE=assign timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -s 3 -c 1 \
                  -o gpurun_out/km_assign python tools/kmeans_probe.py > gpurun_out/ncu_km.log 2>&1
                python tools/ncu_summary.py gpurun_out/km_assign.ncu-rep gpurun_out/km_assign_summary.json > gpurun_out/km_assign_summary.txt 2>&1
                ncu -i gpurun_out/km_assign.ncu-rep --page source --csv > gpurun_out/km_assign_source.csv 2>/dev/null
                rm -f gpurun_out/km_assign.ncu-rep ;;
    kmeans)     timeout 300 python tools/kmeans_probe.py > gpurun_out/kmeans_probe.log 2>&1 ;;
    svg2_launches) timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
                  --log-file gpurun_out/launches_svg2_step.csv python tools/svg2_step_launches.py > gpurun_out/svg2_launches.log 2>&1 ;;
    multi)      # under `gpurun --gpus N`: head-parallel equality test (2 ranks) + bench at N ranks
                N=$(nvidia-smi -L | wc -l)
                timeout 600 python -m pytest tests/test_configs_gpu.py -q -m gpu -k head_parallel --no-header -p no:cacheprovider > gpurun_out/multi_tests.log 2>&1
                timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
                  bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ;;
    *)          if [ -f "$step" ]; then timeout 1200 bash "$step" > "gpurun_out/$(basename "$step").log" 2>&1; else echo "unknown step $step"; fi ;;
  esac
  echo "$step rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/session.log
done
tail -n 3 gpurun_out/*.log 2>/dev/null | tail -n 60
