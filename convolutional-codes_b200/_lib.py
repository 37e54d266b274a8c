"""ctypes binding of lib/libvdec.so (include/vdec.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.environ.get("VDEC_LIB_DIR") or os.path.join(HERE, "lib")      # override: A/B tests of builds


class VdecConfig(C.Structure):
    _fields_ = [("symlen_out", C.c_uint32), ("constr_len", C.c_uint32), ("block_len", C.c_uint32),
                ("metric", C.c_uint32), ("parity", C.c_uint32), ("flags", C.c_uint32),
                ("device", C.c_int32), ("max_frames", C.c_uint32), ("polynomials", C.c_uint64 * 8),
                ("metric_max", C.c_uint32), ("reserved", C.c_uint32)]


class VdecSweepConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("frames_per_batch", C.c_uint64), ("max_frames", C.c_uint64),
                ("min_frame_errors", C.c_uint64), ("chunk_frames", C.c_uint32), ("channel", C.c_uint32),
                ("resume_path", C.c_char_p)]


EXPORTS = {
    # name: (restype, argtypes)
    "vdec_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VdecConfig)]),
    "vdec_destroy": (None, [C.c_void_p]),
    "vdec_last_error": (C.c_char_p, []),
    "vdec_config_labels": (C.c_int, [C.POINTER(VdecConfig), C.c_void_p]),
    "vdec_config_kernel_name": (C.c_char_p, [C.POINTER(VdecConfig)]),
    "vdec_jit_check": (C.c_long, [C.POINTER(VdecConfig)]),
    "vdec_jit_check_iq": (C.c_long, [C.POINTER(VdecConfig)]),
    "vdec_steps": (C.c_uint32, [C.c_void_p]),
    "vdec_states": (C.c_uint32, [C.c_void_p]),
    "vdec_in_bytes_per_frame": (C.c_size_t, [C.c_void_p]),
    "vdec_out_bytes_per_frame": (C.c_size_t, [C.c_void_p]),
    "vdec_kernel_name": (C.c_char_p, [C.c_void_p]),
    "vdec_labels": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vdec_decode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vdec_decode_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vdec_decode_iq_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vdec_decode_iq_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "vdec_stream_decode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_void_p, C.c_void_p]),
    "vdec_sync": (C.c_int, [C.c_void_p]),
    "vdec_debug_check": (C.c_long, [C.c_void_p]),
    "vdec_launch_count": (C.c_uint64, [C.c_void_p]),
    "vdec_count_errors_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vdec_demap_device": (C.c_int, [C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vdec_synth_awgn_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "vdec_hard_demap_device": (C.c_int, [C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vdec_iq_kernel_name": (C.c_char_p, [C.c_void_p]),
    "vdec_sweep": (C.c_int, [C.POINTER(VdecConfig), C.c_void_p, C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_int), C.c_int,
                             C.POINTER(C.c_uint64)]),
    "vdec_device_count": (C.c_int, []),
    "vdec_synth_bsc_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
}

LEGACY_EXPORTS = ("decoder_create", "decoder_destroy", "decoder_init", "decoder_register_callback",
                  "decoder_reset", "decoder_input")

_lib = None


def path(name: str = "libvdec.so") -> str:
    return os.path.join(LIBDIR, name)


def load() -> C.CDLL:
    """Load libvdec.so; raises if it was not built (there is no fallback implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    p = path()
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU or PyTorch fallback for the decoder.")
    lib = C.CDLL(p)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().vdec_last_error() or b"").decode()
