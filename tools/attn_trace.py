"""Dump the per-chunk timeline of one attention CTA (needs a -DSVGB_ATTN_TRACE=<item> build: see build.py --variant).

    python sparse-videogen_b200/build.py --variant trace0 -DSVGB_ATTN_TRACE=0
    SVGB200_LIB=$PWD/sparse-videogen_b200/svgb200/_lib/libsvgb200_trace0.so TRACE_CASE=vb python tools/attn_trace.py

TRACE_CASE: band (HY band plan, two-tile per-tile mapping) | aligned (single-tile items, full chunks) |
            vb (QC=400 / KC=1000 uniform map: item 0 = 256-row two-tile item, item 1 = 41-row single-tile tail; shared
            softmax mapping).  Output: gpurun_out/attn_trace_<case>_<tag>.json + a summary on stdout."""
import ctypes as C
import json
import os
import statistics as st
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from svgb200 import _lib, core  # noqa: E402

dev = torch.device("cuda:0")
H, S, D = 4, bench.S, bench.D
q, k, v = (torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
CASE = os.environ.get("TRACE_CASE", "band")
TAG = os.environ.get("TRACE_TAG", "")
if CASE == "band":
    W, _ = bench.band_width()
    plan = core.plan_band(core.MASK_HY, bench.F * bench.P, bench.F * bench.P + bench.PROMPT_LEN, W, H, S, dev)
elif CASE == "vb":
    bm, row, col, _ = bench.svg2_map(H)
    plan = core.plan_varblock(bm.to(dev), row.to(dev), col.to(dev), S)
else:  # aligned single-tile items, full 128-column chunks
    S = 128 * 930
    q, k, v = (x[:, :, :S].contiguous() for x in (q, k, v))
    g = torch.Generator().manual_seed(0)
    sz = torch.full((H, 930), 128, dtype=torch.int32)
    bm = torch.rand(H, 930, 930, generator=g) < 0.3
    bm[:, :, 0] = True
    plan = core.plan_varblock(bm.to(dev), sz.to(dev), sz.to(dev), S)
for _ in range(3):
    core.attn_fwd(q, k, v, plan)
torch.cuda.synchronize()
buf = (C.c_longlong * 1536)()
fn = _lib.lib().svgb_debug_attn_trace
fn.argtypes = [C.POINTER(C.c_longlong), C.c_int]
assert fn(buf, 1536) == 0
tr = [[[buf[(r * 64 + j) * 8 + e] for e in range(8)] for j in range(64)] for r in range(3)]
t0 = min([x for x in tr[0][2] if x > 0] or [0])
rows = [{"j": j, "w4": [x - t0 for x in tr[0][j][:8]], "w8": [x - t0 for x in tr[1][j][:8]],
         "mma": [x - t0 for x in tr[2][j][:8]]} for j in range(2, 60)]
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"attn_trace_{CASE}_{TAG}.json").write_text(json.dumps(
    {"case": CASE, "tag": TAG, "softmax_events": ["wait_S", "S_ready", "ld_done", "max(+exchange)_done", "exp_done", "st_done", "arrived"],
     "rows": rows}))


def d(role, a, b, lo=6, hi=50):
    xs = [r[role][b] - r[role][a] for r in rows[lo:hi] if r[role][a] > 0 and r[role][b] > 0]
    return round(st.mean(xs), 1) if xs else None


def period(role, ev, lo=6, hi=50):
    xs = [rows[i + 1][role][ev] - rows[i][role][ev] for i in range(lo, hi) if rows[i][role][ev] > 0 and rows[i + 1][role][ev] > 0]
    return round(st.mean(xs), 1) if xs else None


for role in ("w4", "w8"):
    print(CASE, TAG, role, "wait", d(role, 0, 1), "ld", d(role, 1, 2), "max", d(role, 2, 3), "exp", d(role, 3, 4), "st", d(role, 4, 5),
          "arrive", d(role, 5, 6), "e6-7", d(role, 6, 7), "| step period", period(role, 0))
print(CASE, TAG, "mma: waitP", d("mma", 0, 1), "issue1", d("mma", 1, 2), "issue2", d("mma", 2, 3), "e3-4", d("mma", 3, 4), "e4-5", d("mma", 4, 5),
      "e5-6", d("mma", 5, 6), "| period", period("mma", 0))
