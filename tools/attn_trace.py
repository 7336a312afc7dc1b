"""Dump the per-chunk timeline of one attention CTA (needs the SVGB_ATTN_TRACE build).  Bring-up tool."""
import ctypes as C
import json
import os
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
if CASE == "band":
    W, _ = bench.band_width()
    plan = core.plan_band(core.MASK_HY, bench.F * bench.P, bench.F * bench.P + bench.PROMPT_LEN, W, H, S, dev)
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
t0 = tr[2][2][0]
out = {"roles": ["softmax_t0(warp4)", "softmax_t1(warp8)", "mma"],
       "softmax_events": ["wait_S", "S_ready", "ld_done", "max_done", "exp_done", "st_done", "arrived"],
       "mma_events": ["wait_P0", "P0_ready", "pv0_issued", "qk0_issued+commit", "P1_ready", "pv1_issued", "qk1_issued+commit"]}
rows = []
for j in range(2, 40):
    rows.append({"j": j, "t0": [x - t0 for x in tr[0][j][:7]], "t1": [x - t0 for x in tr[1][j][:7]],
                 "mma": [x - t0 for x in tr[2][j][:7]]})
out["rows"] = rows
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "attn_trace.json").write_text(json.dumps(out))
for r in rows[8:12]:
    print(CASE, r["j"], "T0", r["t0"], "T1", r["t1"], "MMA", r["mma"])
# summary: mean durations
import statistics as st
def d(role, a, b):
    return st.mean(r[role][b] - r[role][a] for r in rows[4:36])
print("softmax t0: wait", d("t0", 0, 1), "ld", d("t0", 1, 2), "max", d("t0", 2, 3), "exp", d("t0", 3, 4), "st", d("t0", 4, 5), "arrive", d("t0", 5, 6))
print("softmax t1: wait", d("t1", 0, 1), "ld", d("t1", 1, 2), "max", d("t1", 2, 3), "exp", d("t1", 3, 4), "st", d("t1", 4, 5), "arrive", d("t1", 5, 6))
if CASE == "band":
    print("mma: waitP0", d("mma", 0, 1), "pv0", d("mma", 1, 2), "qk0", d("mma", 2, 3), "waitP1", d("mma", 3, 4), "pv1", d("mma", 4, 5), "qk1", d("mma", 5, 6))
else:
    print("mma: waitP0", d("mma", 0, 1), "pv0", d("mma", 1, 2), "qk0", d("mma", 2, 3))
print("period", st.mean(rows[i + 1]["mma"][0] - rows[i]["mma"][0] for i in range(4, 34)))
