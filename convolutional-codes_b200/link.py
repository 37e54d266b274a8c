"""Frame-sharded link simulation on top of the decoder.

Frames are independent, so ranks take contiguous ranges of the global frame index and the only
exchange is a sum of the three error counters (bits, bit errors, frame errors) — the quantities the
reference drivers accumulate in ``data_callback_decoded`` (AWGN-channel/main.c:129-141) and print
per operating point (``:233``).  Synthetic data is keyed by the global frame index, so totals do
not depend on the number of ranks.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) of `total` frames for `rank`; sizes differ by at most one."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allreduce_counters(counters):
    """Sum int64[3] counters over ranks (NCCL on GPU tensors, gloo on CPU); no-op without a group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def simulate_awgn_point(dec, total_frames: int, ebn0_db: float, seed: int, rank: int = 0, world: int = 1,
                        chunk: int = 1 << 16):
    """One Eb/N0 point of the AWGN driver's loop (AWGN-channel/main.c:193-234), sharded over ranks.
    Returns a CUDA int64[3] tensor [bits, bit errors, frame errors] summed over all ranks."""
    import torch
    from .codes import awgn_sigma
    first, count = shard_range(total_frames, rank, world)
    counters = torch.zeros(3, dtype=torch.int64, device=torch.device("cuda", dec.device))
    done = 0
    while done < count:
        n = min(chunk, count - done)
        payload, metrics = dec.synth_awgn(n, first + done, seed, awgn_sigma(ebn0_db))
        out = dec.decode(metrics)
        dec.count_errors(out, payload, counters)
        done += n
    return allreduce_counters(counters)
