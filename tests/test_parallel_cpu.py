"""world_size-2 gloo test of the head-parallel sharding + output all-gather plumbing.
The per-head compute is the CPU oracle here (tests only); on the GPU the same class wraps the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "sparse-videogen_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.attention import masked_attention_bhsd, wan_mask_mod
    from svgb200.parallel import HeadParallel, owned_heads, shard_heads

    H, S, D = 6, 200, 32
    g = torch.Generator().manual_seed(0)  # same inputs on every rank
    q, k, v = (torch.randn(1, H, S, D, generator=g) for _ in range(3))
    hp = HeadParallel()
    assert hp.local_heads(H) == owned_heads(H, world, rank) == list(range(rank, H, world))
    ql, kl, vl = (shard_heads(t, world, rank) for t in (q, k, v))
    mod = wan_mask_mod(4, 50, 1.0)
    o_local = masked_attention_bhsd(ql[0], kl[0], vl[0], mod)[None]
    full = hp.gather_heads(o_local)
    ref = masked_attention_bhsd(q[0], k[0], v[0], mod)[None]
    ok = torch.allclose(full, ref, atol=1e-6)
    ret[rank] = bool(ok) and tuple(full.shape) == (1, H, S, D)
    dist.destroy_process_group()


def test_head_parallel_gather_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)
