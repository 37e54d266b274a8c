"""Host-side mirror of the reference's decoder interface, over the C ABI.

``BatchDecoder``  — the batched entry points of include/vdec.h (torch CUDA tensors or numpy host
                    arrays in, same formats as the reference decoder's input/output).
``LegacyDecoder`` — the six-function plugin object of ``decoder.h`` (create / init /
                    register_callback / reset / input / destroy) loaded from the shim libraries
                    exactly as the reference's ``main.c`` would link them.

PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib
from .codes import CodeParam, HARD_U16, PARITY_REF, SOFT_F32, SOFT_U8


class VdecError(RuntimeError):
    pass


def make_config(code: CodeParam, metric: int, parity: int, device: int = 0, flags: int = 0,
                max_frames: int = 0, metric_max: int = 0) -> _lib.VdecConfig:
    cfg = _lib.VdecConfig()
    cfg.symlen_out, cfg.constr_len, cfg.block_len = code.symlen_out, code.constr_len, code.block_len
    cfg.metric, cfg.parity, cfg.flags, cfg.device, cfg.max_frames = metric, parity, flags, device, max_frames
    for i, p in enumerate(code.polynomials):
        cfg.polynomials[i] = p
    cfg.metric_max = metric_max
    return cfg


def config_labels(code: CodeParam, parity: int = PARITY_REF) -> np.ndarray:
    """Trellis labels [S][2] as the library computes them on the host (no GPU needed)."""
    lib = _lib.load()
    cfg = make_config(code, SOFT_F32, parity)
    out = np.zeros((code.S, 2), np.uint8)
    if lib.vdec_config_labels(C.byref(cfg), out.ctypes.data) != 0:
        raise VdecError(_lib.last_error())
    return out


def config_kernel_name(code: CodeParam, metric: int, parity: int = PARITY_REF, flags: int = 0, metric_max: int = 0) -> str:
    lib = _lib.load()
    cfg = make_config(code, metric, parity, flags=flags, metric_max=metric_max)
    return lib.vdec_config_kernel_name(C.byref(cfg)).decode()


def jit_check(code: CodeParam, metric: int, parity: int = PARITY_REF, flags: int = 0, metric_max: int = 0, iq: bool = False) -> int:
    """Compile (NVRTC, no GPU needed) the kernel an unregistered code would get; cubin bytes, 0 = nothing to compile.
    iq=True: the raw-I/Q variant of the register kernel (demapper in the feeder)."""
    lib = _lib.load()
    cfg = make_config(code, metric, parity, flags=flags, metric_max=metric_max)
    n = (lib.vdec_jit_check_iq if iq else lib.vdec_jit_check)(C.byref(cfg))
    if n < 0:
        raise VdecError(_lib.last_error())
    return int(n)


class BatchDecoder:
    """One decoder object per (code, metric, parity, device) — ``decoder_create`` + ``decoder_init``."""

    def __init__(self, code: CodeParam, metric: int = SOFT_F32, parity: int = PARITY_REF, device: int = 0,
                 flags: int = 0, max_frames: int = 0, metric_max: int = 0):
        self.lib = _lib.load()
        self.code, self.metric, self.parity, self.device = code, metric, parity, device
        cfg = make_config(code, metric, parity, device, flags, max_frames, metric_max)
        h = C.c_void_p()
        if self.lib.vdec_create(C.byref(h), C.byref(cfg)) != 0:
            raise VdecError(_lib.last_error())
        self.h = h
        self.T = self.lib.vdec_steps(h)
        self.in_bytes = self.lib.vdec_in_bytes_per_frame(h)
        self.out_bytes = self.lib.vdec_out_bytes_per_frame(h)
        self.kernel_name = self.lib.vdec_kernel_name(h).decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.vdec_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def launches(self) -> int:
        return int(self.lib.vdec_launch_count(self.h))

    def labels(self) -> np.ndarray:
        out = np.zeros((self.code.S, 2), np.uint8)
        if self.lib.vdec_labels(self.h, out.ctypes.data) != 0:
            raise VdecError(_lib.last_error())
        return out

    # -- device buffers (torch tensors on cuda:<device>) ----------------------------------------
    def _stream_ptr(self, stream):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    def decode(self, x, out=None, path_metric=None, stream=None):
        """x: soft float32 [F,T,M] or hard uint8 [F,T], CUDA, contiguous -> uint8 [F,(T+7)/8]."""
        import torch
        assert x.is_cuda and x.is_contiguous()
        assert x.dtype == (torch.float32 if self.metric == SOFT_F32 else torch.uint8)
        F = x.shape[0]
        assert x.numel() * x.element_size() == F * self.in_bytes, "input shape does not match the code"
        if out is None:
            out = torch.empty((F, self.out_bytes), dtype=torch.uint8, device=x.device)
        pm_ptr = None
        if self.metric == HARD_U16:
            if path_metric is None:
                path_metric = torch.empty((F,), dtype=torch.int32, device=x.device)
            pm_ptr = C.c_void_p(path_metric.data_ptr())
        rc = self.lib.vdec_decode_device(self.h, C.c_void_p(x.data_ptr()), F, C.c_void_p(out.data_ptr()), pm_ptr,
                                         self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        return (out, path_metric) if self.metric == HARD_U16 else out

    # -- host buffers (numpy or pinned torch CPU tensors): H2D + kernels + D2H inside the call ------
    def decode_host(self, x, out=None, path_metric=None):
        def ptr(a):
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, np.float32 if self.metric == SOFT_F32 else np.uint8)
            nbytes = x.nbytes
        else:
            assert not x.is_cuda and x.is_contiguous()
            nbytes = x.numel() * x.element_size()
        F = nbytes // self.in_bytes
        assert F * self.in_bytes == nbytes, "input size is not a whole number of frames"
        if out is None:
            out = np.empty((F, self.out_bytes), np.uint8)
        pm_ptr = None
        if self.metric == HARD_U16:
            if path_metric is None:
                path_metric = np.empty((F,), np.int32)
            pm_ptr = C.c_void_p(ptr(path_metric))
        rc = self.lib.vdec_decode_host(self.h, C.c_void_p(ptr(x)), F, C.c_void_p(ptr(out)), pm_ptr)
        if rc != 0:
            raise VdecError(_lib.last_error())
        return (out, path_metric) if self.metric == HARD_U16 else out

    # -- raw-I/Q entry: demapper + decoder in one call (common/demapper.c:61-85 + viterbi-decoder.c) --------
    def decode_iq(self, iq, out=None, stream=None):
        """iq: float32 [F,T,2] CUDA, the channel output per trellis step -> uint8 [F,(T+7)/8]."""
        import torch
        assert iq.is_cuda and iq.is_contiguous() and iq.dtype == torch.float32 and iq.shape[1:] == (self.T, 2)
        F = iq.shape[0]
        if out is None:
            out = torch.empty((F, self.out_bytes), dtype=torch.uint8, device=iq.device)
        rc = self.lib.vdec_decode_iq_device(self.h, C.c_void_p(iq.data_ptr()), F, C.c_void_p(out.data_ptr()),
                                            self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        return out

    def decode_iq_host(self, iq, out=None):
        def ptr(a):
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, np.float32)
            n = iq.size
        else:
            assert not iq.is_cuda and iq.is_contiguous()
            n = iq.numel()
        F = n // (self.T * 2)
        assert F * self.T * 2 == n
        if out is None:
            out = np.empty((F, self.out_bytes), np.uint8)
        if self.lib.vdec_decode_iq_host(self.h, C.c_void_p(ptr(iq)), F, C.c_void_p(ptr(out))) != 0:
            raise VdecError(_lib.last_error())
        return out

    # -- sliding-window decoding of one long / unterminated stream -----------------------------------------
    def decode_stream(self, x, block_steps: int = 1024, depth: int = 128, start_known: bool = True, stream=None):
        """x: [nsteps, M] float32 (soft), [nsteps, M] uint8 (quantised) or [nsteps] uint8 (hard), CUDA.
        Returns uint8 [(nsteps+7)//8]: decoded bit per trellis step, MSB-first."""
        import torch
        assert x.is_cuda and x.is_contiguous()
        nsteps = x.shape[0]
        out = torch.empty(((nsteps + 7) // 8,), dtype=torch.uint8, device=x.device)
        rc = self.lib.vdec_stream_decode_device(self.h, C.c_void_p(x.data_ptr()), nsteps, block_steps, depth,
                                                1 if start_known else 0, C.c_void_p(out.data_ptr()), self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        return out

    def sync(self):
        if self.lib.vdec_sync(self.h) != 0:
            raise VdecError(_lib.last_error())

    def debug_check(self) -> int:
        """guard bytes around the library's own device buffers that were overwritten (handle created under VDEC_GUARD=1)"""
        n = int(self.lib.vdec_debug_check(self.h))
        if n < 0:
            raise VdecError(_lib.last_error())
        return n

    # -- the stages either side of the decoder ----------------------------------------------------
    def synth_awgn(self, nframes: int, first_frame: int, seed: int, sigma: float, want_iq: bool = False, stream=None,
                   want_metrics: bool = True):
        """GPU-resident source -> encoder -> mapper -> AWGN [-> soft demapper].  Returns (payload, metrics),
        (payload, metrics, iq) with want_iq, or (payload, iq) with want_metrics=False (the cheap form, for decode_iq)."""
        import torch
        assert want_metrics or want_iq
        dev = torch.device("cuda", self.device)
        c = self.code
        payload = torch.empty((nframes, c.payload_bytes), dtype=torch.uint8, device=dev)
        metrics = torch.empty((nframes, c.T, c.M), dtype=torch.float32, device=dev) if want_metrics else None
        iq = torch.empty((nframes, c.T, 2), dtype=torch.float32, device=dev) if want_iq else None
        rc = self.lib.vdec_synth_awgn_device(self.h, nframes, first_frame, seed, sigma, C.c_void_p(payload.data_ptr()),
                                             C.c_void_p(metrics.data_ptr()) if want_metrics else None,
                                             C.c_void_p(iq.data_ptr()) if want_iq else None, self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        if not want_metrics:
            return payload, iq
        return (payload, metrics, iq) if want_iq else (payload, metrics)

    @property
    def iq_kernel_name(self) -> str:
        """kernels behind decode_iq: fused (demapper in the decoder's feeder) or demap kernel + decoder"""
        return (self.lib.vdec_iq_kernel_name(self.h) or b"").decode()

    def synth_bsc(self, nframes: int, first_frame: int, seed: int, crossover_ppm: int, stream=None):
        import torch
        dev = torch.device("cuda", self.device)
        c = self.code
        payload = torch.empty((nframes, c.payload_bytes), dtype=torch.uint8, device=dev)
        symbols = torch.empty((nframes, c.T), dtype=torch.uint8, device=dev)
        rc = self.lib.vdec_synth_bsc_device(self.h, nframes, first_frame, seed, crossover_ppm,
                                            C.c_void_p(payload.data_ptr()), C.c_void_p(symbols.data_ptr()),
                                            self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        return payload, symbols

    def count_errors(self, decoded, payload, counters=None, stream=None):
        """counters (int64[3], cuda) += [bits, bit errors, frame errors] (AWGN-channel/main.c:129-141)."""
        import torch
        if counters is None:
            counters = torch.zeros(3, dtype=torch.int64, device=decoded.device)
        rc = self.lib.vdec_count_errors_device(self.h, C.c_void_p(decoded.data_ptr()), C.c_void_p(payload.data_ptr()),
                                               decoded.shape[0], C.c_void_p(counters.data_ptr()), self._stream_ptr(stream))
        if rc != 0:
            raise VdecError(_lib.last_error())
        return counters


def demap(n: int, iq, stream=None, hard: bool = False):
    """Demapper on the GPU: soft (bit-exact to common/demapper.c:66-70) or, with hard=True, the hard demapper
    (common/hard-demapper.c:62-102: snap to the nearest point, then the same expression).  iq: float32 [...,2] CUDA."""
    import torch
    lib = _lib.load()
    assert iq.is_cuda and iq.is_contiguous() and iq.dtype == torch.float32
    nsym = iq.numel() // 2
    out = torch.empty(iq.shape[:-1] + (1 << n,), dtype=torch.float32, device=iq.device)
    s = stream if stream is not None else torch.cuda.current_stream(iq.device)
    fn = lib.vdec_hard_demap_device if hard else lib.vdec_demap_device
    if fn(n, C.c_void_p(iq.data_ptr()), nsym, C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream)) != 0:
        raise VdecError(_lib.last_error())
    return out


CHANNEL_AWGN, CHANNEL_AWGN_HARD, CHANNEL_BSC = 0, 1, 2


def sweep(code: CodeParam, points, frames_per_batch: int, channel: int = CHANNEL_AWGN, parity: int = PARITY_REF, seed: int = 12345,
          ndev: int = 0, devices=None, min_frame_errors: int = 0, max_frames: int = 0, chunk_frames: int = 0,
          resume_path: Optional[str] = None, flags: int = 0):
    """The C entry ``vdec_sweep``: the drivers' operating-point loop (AWGN-channel/main.c:193-234,
    binary-symmetric-channel/main.c:147-176) on `ndev` GPUs of this node from ONE process (0 = all), counters summed
    with ncclAllReduce.  points: Eb/N0 in dB, or crossover rates in ppm for CHANNEL_BSC.
    Returns an int array [len(points)][3] = bits, bit errors, frame errors."""
    lib = _lib.load()
    cfg = make_config(code, HARD_U16 if channel == CHANNEL_BSC else SOFT_F32, parity, flags=flags)
    sw = _lib.VdecSweepConfig(seed, frames_per_batch, max_frames, min_frame_errors, chunk_frames, channel,
                              resume_path.encode() if resume_path else None)
    pts = (C.c_double * len(points))(*[float(p) for p in points])
    out = np.zeros((len(points), 3), np.uint64)
    devs = (C.c_int * len(devices))(*devices) if devices else None
    rc = lib.vdec_sweep(C.byref(cfg), C.byref(sw), pts, len(points), devs, len(devices) if devices else ndev,
                        out.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise VdecError(_lib.last_error())
    return out


# ---- legacy plugin object --------------------------------------------------------------------------
class _CodeParamC(C.Structure):     # common/include/code.h:9-19
    _fields_ = [("symlen_out", C.c_uint8), ("constr_len", C.c_uint8), ("block_len", C.c_uint8),
                ("polynomials", C.POINTER(C.c_uint64)), ("bit_metrics", C.c_void_p), ("fano_bit_metrics", C.c_void_p),
                ("metric_weight", C.c_float), ("fano_metric_weight", C.c_float), ("userdata", C.c_void_p)]


_CB_AWGN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_uint8, C.c_void_p)
_CB_BSC = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_uint8, C.c_int32, C.c_void_p)


class LegacyDecoder:
    """The reference's ``struct decoder`` API served by libvdec_awgn.so / libvdec_bsc.so.

    ``kind`` is "awgn" (``decoder_input(float*, len)``) or "bsc" (``decoder_input(uint8_t*, len)``).
    ``lib_path`` may point at any library exporting the six symbols — tests pass the compiled
    reference object to run the same script against both.
    """

    def __init__(self, kind: str, lib_path: Optional[str] = None):
        assert kind in ("awgn", "bsc")
        self.kind = kind
        self.lib = C.CDLL(lib_path or _lib.path(f"libvdec_{kind}.so"))
        L = self.lib
        L.decoder_create.restype = C.c_void_p
        L.decoder_create.argtypes = []
        L.decoder_destroy.argtypes = [C.POINTER(C.c_void_p)]
        L.decoder_destroy.restype = None
        L.decoder_init.argtypes = [C.c_void_p, C.POINTER(_CodeParamC)]
        L.decoder_init.restype = C.c_int
        L.decoder_register_callback.argtypes = [C.c_void_p, C.c_void_p]
        L.decoder_register_callback.restype = None
        L.decoder_reset.argtypes = [C.c_void_p]
        L.decoder_reset.restype = None
        L.decoder_input.argtypes = [C.c_void_p, C.c_void_p, C.c_uint8]
        L.decoder_input.restype = C.c_int
        self.obj = C.c_void_p(L.decoder_create())
        self._cb = None
        self._polys = None
        self.frames = []          # (bytes, len[, path_metric]) captured by the default callback
        self.out_bytes = 0

    def init(self, code: CodeParam) -> int:
        self._polys = (C.c_uint64 * len(code.polynomials))(*code.polynomials)
        self._param = _CodeParamC(code.symlen_out, code.constr_len, code.block_len, self._polys, None, None, 0.0, 0.0, None)
        self.out_bytes = code.out_bytes
        return self.lib.decoder_init(self.obj, C.byref(self._param))

    def register_callback(self, fn: Optional[Callable] = None, ret: int = -1):
        """fn(data: bytes, len, [path_metric]) is called synchronously; returns `ret` to the decoder."""
        ob = lambda: self.out_bytes
        if self.kind == "awgn":
            def cb(data, ln, ud):
                buf = bytes(C.cast(data, C.POINTER(C.c_uint8 * ob())).contents)
                (fn or (lambda *a: self.frames.append(a)))(buf, ln)
                return ret
            self._cb = _CB_AWGN(cb)
        else:
            def cb(data, ln, pm, ud):
                buf = bytes(C.cast(data, C.POINTER(C.c_uint8 * ob())).contents)
                (fn or (lambda *a: self.frames.append(a)))(buf, ln, pm)
                return ret
            self._cb = _CB_BSC(cb)
        self.lib.decoder_register_callback(self.obj, C.cast(self._cb, C.c_void_p))

    def reset(self):
        self.lib.decoder_reset(self.obj)

    def input(self, data: np.ndarray) -> int:
        a = np.ascontiguousarray(data, np.float32 if self.kind == "awgn" else np.uint8)
        assert a.size <= 255
        return self.lib.decoder_input(self.obj, a.ctypes.data, a.size)

    def destroy(self):
        if self.obj:
            self.lib.decoder_destroy(C.byref(self.obj))

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
