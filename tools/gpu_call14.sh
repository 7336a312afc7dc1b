#!/bin/bash
cd "$(dirname "$0")/.."
echo "== k-means tests + probe"
timeout 600 python -m pytest tests/test_svg2_ops_gpu.py tests/test_reference_golden_gpu.py -q -m gpu -rf --no-header -p no:cacheprovider 2>&1 | tail -15
timeout 300 python tools/kmeans_probe.py 2>&1 | tail -12
echo "== synccheck"
timeout 420 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest tests/test_attention_gpu.py -q -m gpu -p no:cacheprovider --no-header \
     -k "selftest_tile or transposed_tail or empty_rows or (band_attention and bfloat16)" > gpurun_out/sanitizer_synccheck.log 2>&1
grep -v "^=========     \|^$" gpurun_out/sanitizer_synccheck.log | grep "=========\|FAILED\|passed\|failed" | head -40
echo "== ncu: band, varblock (main kernel), tail kernel"
PROFILE_MODE=band PROFILE_H=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/attn_band python tools/profile_attn.py > gpurun_out/ncu_band.log 2>&1
PROFILE_MODE=varblock PROFILE_QCKC=400,1000 PROFILE_H=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/attn_varblock python tools/profile_attn.py > gpurun_out/ncu_vb.log 2>&1
PROFILE_MODE=varblock PROFILE_QCKC=400,1000 PROFILE_H=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tail_kernel -c 1 -o gpurun_out/attn_tail python tools/profile_attn.py > gpurun_out/ncu_tail.log 2>&1
for n in band varblock tail; do
  python tools/ncu_summary.py gpurun_out/attn_$n.ncu-rep gpurun_out/attn_${n}_summary.json > gpurun_out/attn_${n}_summary.txt 2>&1
  ls -la gpurun_out/attn_$n.ncu-rep
  rm -f gpurun_out/attn_$n.ncu-rep      # the reports (3 x ~25 MB) do not fit gpurun's 64 MiB return limit
done
grep -E "time_duration|tensor_cycles_active_realtime.avg.pct|dram__bytes_read|dram__bytes_write|pipe_xu|issue_active" gpurun_out/attn_*_summary.txt
echo "== launch list of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_bench.log 2>&1
tail -c 300 gpurun_out/launches_bench.log
du -sh gpurun_out
