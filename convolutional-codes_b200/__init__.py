"""B200-native batched Viterbi decoder behind the C call surface of thomas-emig/convolutional-codes.

The product is ``lib/libvdec.so`` (hand-written sm_100a CUDA + a plain C ABI, ``include/vdec.h``) and
the two legacy shim libraries; this package is the thin host-side mirror used by tests and bench.
"""
from . import codes
from .codes import (CodeParam, FORCE_GENERIC, PREFER_LANEGROUP, PREFER_32BIT, NO_JIT, HARD_U16, PARITY_LINEAR, PARITY_REF, SOFT_F32, SOFT_U8, awgn_sigma, get_code)
from .decoder import (BatchDecoder, LegacyDecoder, VdecError, config_kernel_name, config_labels, demap, jit_check, sweep,
                      CHANNEL_AWGN, CHANNEL_AWGN_HARD, CHANNEL_BSC)
from .link import (BSC_RATES_PPM, allreduce_counters, ber_with_ci, shard_range, simulate_awgn_point, simulate_bsc_point, sweep_awgn,
                   sweep_bsc, write_octave)

__all__ = ["codes", "CodeParam", "get_code", "awgn_sigma", "BatchDecoder", "LegacyDecoder", "VdecError",
           "config_labels", "config_kernel_name", "jit_check", "demap", "shard_range", "allreduce_counters",
           "simulate_awgn_point", "simulate_bsc_point", "sweep_awgn", "sweep_bsc", "BSC_RATES_PPM", "sweep", "CHANNEL_AWGN", "CHANNEL_AWGN_HARD", "CHANNEL_BSC", "write_octave", "ber_with_ci", "PARITY_REF", "PARITY_LINEAR", "SOFT_F32", "SOFT_U8", "HARD_U16", "FORCE_GENERIC", "PREFER_LANEGROUP", "PREFER_32BIT", "NO_JIT"]
