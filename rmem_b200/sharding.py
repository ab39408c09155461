"""Multi-GPU plumbing of the RMem path: independent clips shard across ranks with NO per-frame collective
(the reference shards per clip with a work queue: aot_plus/tools/eval.py:137-145,
aot_plus/networks/managers/evaluator.py:276-295, 589-613).  Collectives exist only at the edges:
one broadcast of the weights at init and one gather of (frames, seconds) at exit.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Static round-robin clip -> rank assignment (clip i -> rank i mod world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_clips, world))


def broadcast_weights(sd: Optional[Dict[str, torch.Tensor]], device, world: int, src: int = 0
                      ) -> Dict[str, torch.Tensor]:
    """Rank `src` holds the state_dict; every rank returns an identical copy.  One metadata object broadcast
    plus ONE flat tensor broadcast (NCCL over NVLink on GPUs, gloo on CPU)."""
    if world <= 1:
        assert sd is not None
        return sd
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = [[(k, tuple(v.shape)) for k, v in sd.items()]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    layout: Sequence[Tuple[str, Tuple[int, ...]]] = meta[0]
    total = sum(int(torch.Size(s).numel()) for _, s in layout)
    if rank == src:
        flat = torch.cat([sd[k].reshape(-1).float() for k, _ in layout]).to(device)
    else:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src)
    out, off = {}, 0
    flat = flat.cpu()
    for k, s in layout:
        n = int(torch.Size(s).numel())
        out[k] = flat[off:off + n].view(*s).clone()
        off += n
    return out


def gather_stats(frames: int, seconds: float, device, world: int) -> List[Tuple[int, float]]:
    """(frames, seconds) of every rank, on every rank."""
    if world <= 1:
        return [(frames, seconds)]
    import torch.distributed as dist
    t = torch.tensor([float(frames), float(seconds)], dtype=torch.float64, device=device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [(int(o[0].item()), float(o[1].item())) for o in outs]
