"""Multi-GPU host logic: streams are independent units (own recurrent state, stateless detector), so they are
partitioned across ranks with no data-path collective; the only collective is one broadcast of the packed weight
blob at init.  Works on any torch.distributed backend (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import List


def shard_streams(n_streams: int, rank: int, world_size: int) -> List[int]:
    """Stream i runs on rank i mod world_size."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return list(range(rank, n_streams, world_size))


def broadcast_blob(blob, src: int = 0) -> None:
    """One broadcast of the packed weights (a uint8 tensor); no-op outside a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
