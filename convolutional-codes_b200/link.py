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


def ber_with_ci(bits: int, bit_errors: int, frame_errors: int):
    """BER and a 95 % interval.  Decoder errors come in bursts, so the independent events are errored
    FRAMES, not bits: relative half-width 1.96/sqrt(frame_errors) (0 errors: upper bound 3/bits)."""
    if bits == 0:
        return 0.0, 0.0, 1.0
    ber = bit_errors / bits
    if frame_errors == 0:
        return 0.0, 0.0, 3.0 / bits
    rel = 1.96 / (frame_errors ** 0.5)
    return ber, max(0.0, ber * (1.0 - rel)), ber * (1.0 + rel)


def _sweep(point_fn, dec, points, frames_per_point, key, seed, rank, world, chunk, min_frame_errors, max_frames):
    out = []
    for i, x in enumerate(points):
        n = frames_per_point[i] if isinstance(frames_per_point, (list, tuple)) else frames_per_point
        total = [0, 0, 0]
        done = 0
        while True:
            c = point_fn(dec, n, x, seed + 7919 * i, rank, world, chunk, first_frame=done)
            c = c.cpu().tolist()
            total = [a + b for a, b in zip(total, c)]
            done += n
            if min_frame_errors <= 0 or total[2] >= min_frame_errors or (max_frames and done >= max_frames):
                break
        ber, lo, hi = ber_with_ci(*total)
        out.append({key: x, "bits": total[0], "bit_errors": total[1], "frame_errors": total[2],
                    "ber": ber, "ci95_lo": lo, "ci95_hi": hi})
    return out


def sweep_awgn(dec, ebn0_db_list, frames_per_point, seed: int = 12345, rank: int = 0, world: int = 1,
               chunk: int = 1 << 16, min_frame_errors: int = 0, max_frames: int = 0, hard_demap: bool = False):
    """The AWGN driver's Eb/N0 loop (AWGN-channel/main.c:193-234) on the GPU, frames sharded over ranks.
    frames_per_point: int or list; with min_frame_errors > 0 a point keeps adding batches of that many
    frames until the (all-rank) frame-error count is reached or max_frames is hit.  hard_demap=True puts
    common/hard-demapper.c in front of the decoder (the `*_h` curves of results/awgn_channel.m:16-28).
    Returns a list of dicts {ebn0_db, bits, bit_errors, frame_errors, ber, ci95_lo, ci95_hi}."""
    def fn(dec, n, x, seed, rank, world, chunk, first_frame):
        return simulate_awgn_point(dec, n, x, seed, rank, world, chunk, first_frame, hard_demap)
    return _sweep(fn, dec, ebn0_db_list, frames_per_point, "ebn0_db", seed, rank, world, chunk, min_frame_errors, max_frames)


def sweep_bsc(dec, crossover_ppm_list, frames_per_point, seed: int = 12345, rank: int = 0, world: int = 1,
              chunk: int = 1 << 16, min_frame_errors: int = 0, max_frames: int = 0):
    """The BSC driver's loop over channel error rates (binary-symmetric-channel/main.c:103-109,147-176) on the
    GPU with the hard-decision decoder; rates in ppm as the driver's table."""
    return _sweep(simulate_bsc_point, dec, crossover_ppm_list, frames_per_point, "crossover_ppm", seed, rank, world, chunk,
                  min_frame_errors, max_frames)


# channel error rates of the reference's BSC driver, in ppm (binary-symmetric-channel/main.c:103-109)
BSC_RATES_PPM = (1, 5, 25, 125, 625, 3125, 6250, 12500, 15625, 25000, 50000, 78125, 100000, 200000, 300000, 390625, 400000)


def write_octave(path: str, name: str, points) -> None:
    """Append a result vector in the style of the reference's results/awgn_channel.m (:1, :14-28)."""
    key = "ebn0_db" if "ebn0_db" in points[0] else "crossover_ppm"
    with open(path, "a") as f:
        f.write("%s_%s = [%s];\n" % ("snr" if key == "ebn0_db" else "ppm", name, " ".join("%g" % p[key] for p in points)))
        f.write("bits_%s = [%s];\n" % (name, " ".join("%d" % p["bits"] for p in points)))
        f.write("ber_%s = [%s];\n" % (name, " ".join("%.6e" % p["ber"] for p in points)))


def simulate_awgn_point(dec, total_frames: int, ebn0_db: float, seed: int, rank: int = 0, world: int = 1,
                        chunk: int = 1 << 16, first_frame: int = 0, hard_demap: bool = False):
    """One Eb/N0 point of the AWGN driver's loop (AWGN-channel/main.c:193-234), sharded over ranks: the source emits
    raw I/Q and the decoder's raw-I/Q entry forms the branch metrics (in the kernel's feeder where the family allows).
    Returns a CUDA int64[3] tensor [bits, bit errors, frame errors] summed over all ranks."""
    import torch
    from .codes import awgn_sigma
    from .decoder import demap
    first, count = shard_range(total_frames, rank, world)
    counters = torch.zeros(3, dtype=torch.int64, device=torch.device("cuda", dec.device))
    done = 0
    while done < count:
        n = min(chunk, count - done)
        payload, iq = dec.synth_awgn(n, first_frame + first + done, seed, awgn_sigma(ebn0_db), want_iq=True, want_metrics=False)
        out = dec.decode(demap(dec.code.n, iq, hard=True)) if hard_demap else dec.decode_iq(iq)
        dec.count_errors(out, payload, counters)
        done += n
    return allreduce_counters(counters)


def simulate_bsc_point(dec, total_frames: int, crossover_ppm: int, seed: int, rank: int = 0, world: int = 1,
                       chunk: int = 1 << 16, first_frame: int = 0):
    """One channel error rate of the BSC driver (binary-symmetric-channel/main.c:57-72,147-176), sharded over ranks."""
    import torch
    first, count = shard_range(total_frames, rank, world)
    counters = torch.zeros(3, dtype=torch.int64, device=torch.device("cuda", dec.device))
    done = 0
    while done < count:
        n = min(chunk, count - done)
        payload, sym = dec.synth_bsc(n, first_frame + first + done, seed, int(crossover_ppm))
        out, _ = dec.decode(sym)
        dec.count_errors(out, payload, counters)
        done += n
    return allreduce_counters(counters)
