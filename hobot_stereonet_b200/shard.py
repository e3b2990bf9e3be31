"""Host-side multi-GPU plumbing: stereo pairs are independent units, so the path scales as
replicas with the batch sharded across ranks.  The only collective is ONE broadcast of the weight blob at
init (the reference loads `model_file` on its single device, stereonet_node.cpp:131-136); nothing is
exchanged on the per-frame path.  Backend: "nccl" on the GPU box, "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `n_pairs` stereo pairs owned by `rank`; the remainder goes to the
    lowest ranks (config 4: 32 pairs over 8 ranks -> 4 each; 10 over 4 -> 3,3,2,2).  Empty shards are legal."""
    if world < 1 or not 0 <= rank < world or n_pairs < 0:
        raise ValueError(f"bad shard request n_pairs={n_pairs} world={world} rank={rank}")
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_blob(blob: Optional[bytes], src: int = 0, device: Optional[torch.device] = None) -> bytes:
    """Rank `src` passes the weight blob (bytes), every other rank passes None; all ranks return the blob.
    Two broadcasts on the wire (an 8-byte length, then the payload) = the single weight collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if blob is None:
            raise ValueError("broadcast_blob: no process group and no blob")
        return bytes(blob)
    rank = dist.get_rank()
    dev = device or torch.device("cpu")
    if rank == src:
        if blob is None:
            raise ValueError("broadcast_blob: the source rank must supply the blob")
        payload = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        n = torch.tensor([payload.numel()], dtype=torch.int64, device=dev)
    else:
        n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank != src:
        payload = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(payload, src)
    return payload.cpu().numpy().tobytes()


def gather_counts(local_pairs: int, local_ms: float, device: Optional[torch.device] = None) -> Tuple[int, float]:
    """Benchmark bookkeeping only: (pairs processed by all ranks, max over ranks of the device time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_pairs, local_ms
    dev = device or torch.device("cpu")
    s = torch.tensor([float(local_pairs)], dtype=torch.float64, device=dev)
    m = torch.tensor([float(local_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(round(float(s.item()))), float(m.item())
