"""Head-parallel execution over the GPUs of one NVSwitch box.

Every op on the hot path is independent per (batch, head), so the only exchange step is the all-gather
of the attention output, and only when the consumer needs all heads.  Rank r owns heads h with
h % world == r (interleaved): the heads gathered for local head index i are then the contiguous slice
[i*world, (i+1)*world) of the full [H, S, D] output, so each per-head all-gather writes straight into
the final tensor (ncclAllGather, no staging copy) and can be issued on a side stream as soon as that
head's attention finishes — the transfer of head i overlaps the attention of head i+1.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist


def owned_heads(num_heads: int, world: int, rank: int):
    """heads owned by `rank` (interleaved assignment)."""
    assert num_heads % world == 0, "number of heads must divide across ranks"
    return list(range(rank, num_heads, world))


def shard_heads(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """[cfg, H, S, D] -> contiguous [cfg, H/world, S, D] slice owned by `rank`."""
    return t[:, rank::world].contiguous()


class HeadParallel:
    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._comm_stream = None

    def streams(self, device, compute_streams: int = 2):
        """(communication stream, compute side streams), created on first use."""
        if self._comm_stream is None:
            # High priority so that NCCL CTAs are scheduled ahead of queued attention CTAs when SMs free up (the attention
            # kernel fills every SM: 1 CTA / SM, the whole register file).  Measured at N = 8 in one session
            # (profiles/r02_hp_priority_ab.txt): 6579 / 6417 TFLOP/s with priority -1, 6430 / 6420 with the default --
            # no measurable difference; what is exposed is the all-gather of the LAST head (~0.9 ms of 8.0), which has
            # no compute left to hide behind.
            prio = int(os.environ.get("SVGB_HP_COMM_PRIORITY", "-1"))  # 0 = default priority (A/B switch)
            self._comm_stream = torch.cuda.Stream(device=device, priority=prio)
            self._compute_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, compute_streams))]
        return self._comm_stream, self._compute_streams

    def local_heads(self, num_heads: int):
        return owned_heads(num_heads, self.world, self.rank)

    def gather_heads(self, o_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """o_local [1, Hl, S, D] -> [1, Hl*world, S, D] (blocking w.r.t. the current stream)."""
        cfg, Hl, S, D = o_local.shape
        assert cfg == 1
        if self.world == 1:
            return o_local
        if out is None:
            out = torch.empty(1, Hl * self.world, S, D, dtype=o_local.dtype, device=o_local.device)
        for i in range(Hl):
            dist.all_gather_into_tensor(out[0, i * self.world:(i + 1) * self.world], o_local[0, i:i + 1].contiguous(),
                                        group=self.group)
        return out

    def run_overlapped(self, per_head_fn: Callable[[int], torch.Tensor], num_local_heads: int, S: int, D: int,
                       dtype, device, out: Optional[torch.Tensor] = None, compute_streams: int = 2,
                       trace: Optional[dict] = None) -> torch.Tensor:
        """per_head_fn(i) computes local head i ([S, D] or [1,1,S,D]) on the CURRENT stream.

        Heads are issued round-robin on `compute_streams` side streams, so the last partial wave of head i's
        attention kernel is back-filled by CTAs of head i+1 (a 467-CTA launch is 3.15 waves on 148 SMs), and
        each head's all-gather runs on a communication stream as soon as that head finished, overlapping the
        following heads' compute (CUDA / NCCL only)."""
        H = num_local_heads * self.world
        if out is None:
            out = torch.empty(1, H, S, D, dtype=dtype, device=device)
        cur = torch.cuda.current_stream(device)
        self.streams(device, compute_streams)
        start = torch.cuda.Event(enable_timing=trace is not None)
        start.record(cur)
        if trace is not None:  # event stamps for the per-rank stage timeline (bench.py `scaling_timeline`)
            trace.update(start=start, compute_done=[], comm_done=[])
        keep = []
        for i in range(num_local_heads):
            cs = self._compute_streams[i % len(self._compute_streams)]
            with torch.cuda.stream(cs):
                cs.wait_event(start)
                o = per_head_fn(i).reshape(1, S, D)
                ev = torch.cuda.Event(enable_timing=trace is not None)
                ev.record(cs)
            if trace is not None:
                trace["compute_done"].append(ev)
            keep.append(o)
            if self.world == 1:
                with torch.cuda.stream(cs):
                    out[0, i] = o[0]
                continue
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                dist.all_gather_into_tensor(out[0, i * self.world:(i + 1) * self.world], o, group=self.group)
                if trace is not None:
                    evc = torch.cuda.Event(enable_timing=True)
                    evc.record(self._comm_stream)
                    trace["comm_done"].append(evc)
        for cs in self._compute_streams:
            cur.wait_stream(cs)
        cur.wait_stream(self._comm_stream)
        for o in keep:  # tensors were produced on side streams: tell the allocator the current stream uses them
            o.record_stream(cur)
        if trace is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(cur)
            trace["end"] = end
        return out
