"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN/bench cite."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
out = {}
for k in hdr:
    for want in keys:
        if k.endswith(want) or k == want:
            out[k] = d[k]
for k, (u, v) in out.items():
    print(f"{k:100s} {v:>18s} {u}")
stalls = {k: float(v[1]) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[1]}
print("-- warp stall reasons (avg warps stalled per issue-active cycle)")
for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
    print(f"   {k.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.3f}")
if len(sys.argv) > 2:
    json.dump({k: {"unit": u, "value": v} for k, (u, v) in out.items()} | {"stalls": stalls}, open(sys.argv[2], "w"), indent=1)
