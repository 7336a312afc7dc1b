#!/usr/bin/env python
"""bench.py — block-sparse attention at HunyuanVideo 720p (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the SVG1 sparse attention core over synthetic Q/K/V of the HunyuanVideo 720p
shape (B=1, H=24, S=119056 = 33 frames x 3600 tokens + 256 text, D=128, bf16) under the executed SVG1
band mask at 30 % density (W = 19072, hyvideo/utils.py:20-44,142-151): online profiling (sample_mse ->
best_mask_idx), head placement of Q/K/V, element-exact block-sparse attention, inverse placement of O —
the sparse branch of attention_core_logic (svg/models/hyvideo/attention.py:506-524).  Heads are sharded
across ranks (strong scaling: total work fixed), the output all-gathered once.

Printed JSON (one line, rank 0): the contract keys plus `roofline` (dominant kernel = the tcgen05
attention kernel, CUDA-event timed on its stream), `cpu_baseline` (oracle naive attention on the host
cores, bounded sample), `e2e` (same step through the public API from pinned host buffers, H2D + D2H in
the timed region), `clocks`, `gpu_launches`, and `svg2` (variable-block kernel at rho = 0.30,
QC=400/KC=1000, the SVG2 shape of the same model — reported, not the headline).

--impl reference: times the reference's own CPU formulation of this path (naive masked
torch attention, svg/kernels/test/test_sparse_attn.py:109-157, restated in oracle/attention.py) on all
host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))

import torch  # noqa: E402

# ---- workload constants (HunyuanVideo T2V 720p) --------------------------------
H_TOTAL, D, CTX, F, P, PROMPT_LEN = 24, 128, 256, 33, 3600, 60
S = CTX + F * P
SPARSITY = 0.30
N_SAMPLED_ROWS, SAMPLE_MAX_ROW = 64, 10000


def band_width(sparsity=SPARSITY):
    """sparsity_to_width (hyvideo/utils.py:142-151) then floor to 128 (:24-25)."""
    seq = S
    s2 = (sparsity * seq * seq - 2 * seq * CTX) / (seq * seq)
    mul = seq * (1 - math.sqrt(1 - s2)) / P
    return math.floor(mul * P / 128) * 128, mul


def band_pairs(W):
    """number of allowed (q, kv) pairs per head under the HY mask_mod (exact, closed form per row)."""
    V, R = F * P, F * P + PROMPT_LEN
    q = torch.arange(S, dtype=torch.int64)
    lo = torch.clamp(q - (W - 1), min=0)
    hi = torch.clamp(q + (W - 1), max=V - 1)
    band = torch.clamp(hi - lo + 1, min=0)
    per_row = torch.where(q < V, band + PROMPT_LEN, torch.where(q < R, torch.tensor(R), torch.tensor(S - R)))
    return int(per_row.sum().item())


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"burst": d.get("bf16_tflops"), "sustained": d.get("bf16_tflops_sustained"), "hbm": d.get("hbm_gbs"),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # under-load median: drop idle samples (below 60 % of the max seen)
        loaded = [x for x in sm if sm and x >= 0.6 * sm[-1]] or sm
        return {"sm_mhz": loaded[len(loaded) // 2] if loaded else None, "sm_max_mhz": mx,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: naive masked attention on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------
def usable_cores():
    """Host threads this process may really use: cpu_count capped by the affinity mask and the cgroup CPU quota
    (a container can report 128 CPUs and be allowed 16; oversubscribing makes the CPU arm look worse than it is)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = (Path("/sys/fs/cgroup/cpu.max").read_text().split() + ["100000"])[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def cpu_naive_sample(rows=1024, seconds=10.0):
    """One head, up to `rows` query rows spread over the sequence (256 at a time, until `seconds` of work are
    done), full S keys, HY band mask, fp32.  Returns (tflops, seconds, sample description, threads).  Same
    arithmetic as ref_torch_attn_impl (svg/kernels/test/test_sparse_attn.py:109-157)."""
    from oracle.attention import hy_mask_mod

    cores = usable_cores()
    torch.set_num_threads(cores)
    W, mul = band_width()
    g = torch.Generator().manual_seed(0)
    k = torch.randn(S, D, generator=g)
    v = torch.randn(S, D, generator=g)
    qrows = torch.linspace(0, S - 1, rows).long()
    qrows = qrows[torch.randperm(rows, generator=g)]  # any prefix of the sample covers the whole sequence
    q = torch.randn(rows, D, generator=g)
    mod = hy_mask_mod(CTX, PROMPT_LEN, F, P, mul)
    kv_idx = torch.arange(S).view(1, S)
    t0 = time.perf_counter()
    pairs, done = 0, 0
    for r0 in range(0, rows, 256):
        qi = qrows[r0:r0 + 256].view(-1, 1)
        s = (q[r0:r0 + 256] @ k.T) / math.sqrt(D)
        m = mod(qi, kv_idx)
        s = s.masked_fill(~m, float("-inf"))
        w = torch.softmax(s, dim=-1)
        _ = w @ v
        pairs += int(m.sum().item())
        done += qi.numel()
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    # the naive formulation computes every (q, kv) pair and masks afterwards; credit only the
    # algorithmic (allowed) pairs so the unit matches the GPU arm
    flops = 4.0 * D * pairs
    return (flops / dt / 1e12, dt,
            f"1 head, {done} query rows spread over S={S}, all keys, HY band W={W}, {cores} threads", cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        tf, dt, sample, cores = cpu_naive_sample(rows=4096, seconds=10.0)
        if i >= args.warmup:
            vals.append(tf)
            secs.append(dt)
    value = sum(vals) / len(vals)
    W, _ = band_width()
    pairs = band_pairs(W)
    ms_call = 4.0 * D * pairs * H_TOTAL / (value * 1e12) * 1e3
    line = {
        "impl": "reference", "metric": "block-sparse attention TFLOP/s (density-adjusted)", "value": value,
        "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sum(secs) / len(secs) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HunyuanVideo-720p SVG1 sparse attention core (sample_mse + placement + band attention "
                               "+ inverse placement), rho=0.30", "S": S, "heads": H_TOTAL, "head_dim": D,
                   "band_W": W, "density": pairs / S / S,
                   "note": "the reference's naive torch attention on the host cores, bounded sample per step; "
                           "ms per call extrapolated", "extrapolated_ms_per_call": ms_call},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from svgb200 import core
    from svgb200.models import hyvideo as hy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    core.device_check()
    assert H_TOTAL % world == 0, "heads must divide across ranks"
    Hl = H_TOTAL // world

    # synthetic inputs: head h is generated from seed + h so results are identical for any N
    def make(seed_off):
        t = torch.empty(1, Hl, S, D, dtype=torch.bfloat16, device=dev)
        for i in range(Hl):
            g = torch.Generator(device=dev).manual_seed(1000 * seed_off + i * world + rank)  # global head id
            t[0, i] = torch.randn(S, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        return t

    q, k, v = make(1), make(2), make(3)
    W, mul = band_width()
    pairs = band_pairs(W)
    flops_total = 4.0 * D * pairs * H_TOTAL
    flops_local = 4.0 * D * pairs * Hl

    proc = hy.HunyuanSVG1Core(context_length=CTX, prompt_length=PROMPT_LEN, num_frame=F, frame_size=P,
                              num_heads=Hl, head_dim=D, sparsity=SPARSITY, num_sampled_rows=N_SAMPLED_ROWS,
                              sample_mse_max_row=SAMPLE_MAX_ROW, device=dev)
    gen = torch.Generator().manual_seed(1234)  # CPU generator like the reference's torch.randint (attention.py:381)
    from svgb200.parallel import HeadParallel

    hp = HeadParallel() if world > 1 else None
    gathered = torch.empty(1, H_TOTAL, S, D, dtype=torch.bfloat16, device=dev) if world > 1 else None
    attn_ev = []

    def step(timed=False):
        rows = torch.randint(0, SAMPLE_MAX_ROW, (N_SAMPLED_ROWS,), generator=gen)
        if world > 1:
            # per-head attention, each head's output all-gather overlapped with the next head's compute
            return proc.sparse_core_head_parallel(q, k, v, hp, sampled_rows=rows, out=gathered,
                                                  attn_events=attn_ev if timed else None)
        return proc.sparse_core(q, k, v, sampled_rows=rows, attn_events=attn_ev if timed else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = core.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(timed=True)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = core.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world == 1:
        attn_ms = sum(a.elapsed_time(b) for a, b in attn_ev) / max(1, args.steps)
    else:
        # per-head launches overlap across streams in the step, so time the dominant kernel on its own:
        # one launch over all local heads (same plan, same inputs), CUDA events on its stream
        core.attn_fwd(q, k, v, proc.block_mask.plan)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            core.attn_fwd(q, k, v, proc.block_mask.plan)
        b.record()
        torch.cuda.synchronize()
        attn_ms = a.elapsed_time(b) / 3
    t = torch.tensor([ms_total, attn_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / args.steps
    attn_ms = t[1].item()

    # ---- e2e: pinned host buffers -> H2D -> step -> D2H of the result, through the public API
    hq, hk, hv = (x.cpu().pin_memory() for x in (q, k, v))
    ho = torch.empty(1, Hl, S, D, dtype=torch.bfloat16).pin_memory()

    def e2e_step():
        rows = torch.randint(0, SAMPLE_MAX_ROW, (N_SAMPLED_ROWS,), generator=gen)
        proc.sparse_core_from_host(hq, hk, hv, ho, sampled_rows=rows)

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    del hq, hk, hv, ho

    # ---- SVG2 variable-block kernel at rho = 0.30 (reported beside the headline)
    svg2 = None
    if rank == 0:
        QC, KC = 400, 1000
        Sv = F * P
        gm = torch.Generator().manual_seed(7)

        def sizes(n, total):
            b = torch.full((Hl, n), total // n, dtype=torch.int32)
            b[:, : total - (total // n) * n] += 1
            return b

        row = torch.cat([sizes(QC, Sv), torch.tensor([[PROMPT_LEN, CTX - PROMPT_LEN]] * Hl, dtype=torch.int32)], 1)
        col = torch.cat([sizes(KC, Sv), torch.tensor([[PROMPT_LEN, CTX - PROMPT_LEN]] * Hl, dtype=torch.int32)], 1)
        bm = torch.zeros(Hl, QC + 2, KC + 2, dtype=torch.bool)
        bm[:, :QC, :KC] = torch.rand(Hl, QC, KC, generator=gm) < 0.30
        bm[:, -2, :-1] = True   # prompt block <-> everything but the padding (attention.py:681-684)
        bm[:, :-1, -2] = True
        bm[:, -1, -1] = True
        fl = 4.0 * D * (row.double()[:, :, None] * col.double()[:, None, :] * bm).sum().item()
        plan = core.plan_varblock(bm.to(dev), row.to(dev), col.to(dev), S)
        for _ in range(2):
            core.attn_fwd(q, k, v, plan)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n_it = 3
        for _ in range(n_it):
            core.attn_fwd(q, k, v, plan)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n_it
        svg2 = {"workload": f"variable-block map QC={QC}+2 KC={KC}+2 Bernoulli(0.30), uniform cluster sizes, {Hl} heads",
                "ms_per_call": ms, "tflops": fl / ms / 1e9, "density": fl / (4.0 * D * Hl * S * S),
                "dense_equiv_tflops": 4.0 * D * Hl * S * S / ms / 1e9}

    # ---- FP8 (e4m3) variant of the same band-mask attention (BASELINE config 5 flavour), reported beside it
    fp8 = None
    if rank == 0:
        try:
            (q8, sq), (k8, sk), (v8, sv) = (core.quantize_e4m3(x) for x in (q, k, v))
            for _ in range(2):
                core.attn_fwd_fp8(q8, k8, v8, sq, sk, sv, proc.block_mask.plan)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                core.attn_fwd_fp8(q8, k8, v8, sq, sk, sv, proc.block_mask.plan)
            b.record()
            torch.cuda.synchronize()
            ms8 = a.elapsed_time(b) / 3
            a.record()
            core.quantize_e4m3(q)
            b.record()
            torch.cuda.synchronize()
            fp8 = {"workload": f"same band mask, e4m3 Q/K/V (per-head scales), bf16 out, {Hl} heads", "ms_per_call": ms8,
                   "tflops": flops_local / ms8 / 1e9, "quantize_ms_per_tensor": a.elapsed_time(b)}
            del q8, k8, v8
        except Exception as e:  # noqa: BLE001
            fp8 = {"error": repr(e)[:200]}

    # ---- pre-attention chain: fused transpose + QK-RMSNorm + RoPE, HBM-bound
    prep = None
    if rank == 0:
        try:
            qi, ki, vi = (torch.randn(1, S, Hl * D, device=dev).bfloat16() for _ in range(3))
            gq, gk = (torch.randn(D, device=dev).bfloat16() for _ in range(2))
            cos, sin = (torch.randn(S - CTX, D, device=dev) for _ in range(2))
            outs = tuple(torch.empty(1, Hl, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))

            def prep_call():
                core.qkv_prep(qi, ki, vi, Hl, out=outs, norm=core.NORM_RMS_HEAD, gamma_q=gq, gamma_k=gk, eps=1e-6, rope=1,
                              cos=cos, sin=sin, rope_lo=0, rope_n=S - CTX)
            for _ in range(2):
                prep_call()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                prep_call()
            b.record()
            torch.cuda.synchronize()
            msp = a.elapsed_time(b) / 5
            nbytes = 6 * Hl * S * D * 2
            hbm = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs") if (ROOT / "MEASURED_PEAKS.json").exists() else None
            prep = {"workload": f"[1,S,{Hl}x{D}] q,k,v -> [1,{Hl},S,{D}]: transpose + per-head RMSNorm(q,k) + RoPE (text last)",
                    "ms_per_call": msp, "algorithmic_bytes": nbytes, "gbs": nbytes / msp / 1e6,
                    "hbm_peak_gbs": hbm, "frac": (nbytes / msp / 1e6 / hbm) if hbm else None}
            del qi, ki, vi, outs
        except Exception as e:  # noqa: BLE001
            prep = {"error": repr(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    achieved = flops_local / attn_ms / 1e9 if attn_ms > 0 else None
    traffic = None
    tj = ROOT / "profiles" / "attn_traffic.json"
    if tj.exists():
        try:
            traffic = json.loads(tj.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    tf_cpu, dt_cpu, sample, cores = cpu_naive_sample(rows=4096, seconds=12.0)
    value = flops_total / ms_step / 1e9
    bytes_in = 3 * Hl * S * D * 2
    bytes_out = Hl * S * D * 2
    line = {
        "metric": "block-sparse attention TFLOP/s (density-adjusted)", "value": value, "unit": "TFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "HunyuanVideo-720p SVG1 sparse attention core (sample_mse + placement + band attention "
                               "+ inverse placement), rho=0.30", "S": S, "heads": H_TOTAL, "head_dim": D,
                   "band_W": W, "density": pairs / S / S, "heads_per_gpu": Hl, "parallelism": f"head-parallel x{world}",
                   "l2": "inputs (2.2 GB/GPU-set) larger than L2; no flush needed",
                   "attn_ms_per_call": attn_ms, "flops_per_call": flops_total},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                     "frac": (achieved / pk["sustained"]) if achieved and pk["sustained"] else None,
                     "frac_of_burst": (achieved / pk["burst"]) if achieved and pk["burst"] else None,
                     "peak_src": pk["src"] + ", sustained (kernel runs ~60 ms back to back)",
                     "kernel": "svgb::attn_fwd_kernel<128,true>", "traffic": traffic},
        "cpu_baseline": {"value": tf_cpu, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                         "seconds": dt_cpu},
        "e2e": {"value": flops_total / e2e_ms / 1e9, "unit": "TFLOP/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out},
        "clocks": clocks, "gpu_launches": launches, "svg2": svg2, "fp8": fp8, "prep": prep,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
