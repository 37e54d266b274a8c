"""TEST INFRASTRUCTURE — NOT A PRODUCT PATH.

ctypes front-end to the CPU oracle (``oracle/vit_oracle.c``) and, when it was
built in a container that has ``/root/reference``, to the UNMODIFIED reference
objects (``oracle/_ref/libref_*.so``, recipe in ``oracle/Makefile``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs import this package.  The product package
(``convolutional-codes_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

PARITY_REF = 0
PARITY_LINEAR = 1


class _OrcCode(C.Structure):
    _fields_ = [("symlen_out", C.c_uint32), ("constr_len", C.c_uint32), ("block_len", C.c_uint32),
                ("parity_mode", C.c_uint32), ("polynomials", C.c_uint64 * 8)]


class _RefhCode(C.Structure):
    _fields_ = [("symlen_out", C.c_uint32), ("constr_len", C.c_uint32), ("block_len", C.c_uint32),
                ("polynomials", C.c_uint64 * 8)]


def build(force: bool = False) -> None:
    """(Re)build oracle/_ref via the Makefile; reference libs only if /root/reference exists."""
    lib = os.path.join(_REF, "libvitoracle.so")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(os.path.join(_HERE, "vit_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Code:
    """Code description shared by oracle and reference wrappers."""

    def __init__(self, n: int, K: int, L: int, polys, parity: int = PARITY_REF):
        self.n, self.K, self.L, self.parity = int(n), int(K), int(L), int(parity)
        self.polys = [int(p) for p in polys]
        assert len(self.polys) == self.n

    @property
    def S(self): return 1 << (self.K - 1)
    @property
    def T(self): return self.L + self.K - 1
    @property
    def M(self): return 1 << self.n
    @property
    def out_bytes(self): return (self.T + 7) // 8
    @property
    def payload_bytes(self): return (self.L + 7) // 8

    def _orc(self):
        c = _OrcCode(self.n, self.K, self.L, self.parity)
        for i, p in enumerate(self.polys):
            c.polynomials[i] = p
        return c

    def _refh(self):
        c = _RefhCode(self.n, self.K, self.L)
        for i, p in enumerate(self.polys):
            c.polynomials[i] = p
        return c


class Oracle:
    """The widened CPU restatement (always available once built)."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_REF, "libvitoracle.so"))
        L = self.lib
        L.orc_parity.restype = C.c_uint32
        L.orc_parity.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_decode_soft.restype = C.c_int
        L.orc_decode_soft.argtypes = [C.POINTER(_OrcCode), C.POINTER(C.c_float), C.c_uint64, C.POINTER(C.c_uint8)]
        L.orc_decode_hard.restype = C.c_int
        L.orc_decode_hard.argtypes = [C.POINTER(_OrcCode), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        L.orc_encode.restype = None
        L.orc_encode.argtypes = [C.POINTER(_OrcCode), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8)]
        L.orc_labels.restype = None
        L.orc_labels.argtypes = [C.POINTER(_OrcCode), C.POINTER(C.c_uint8)]
        for fn in (L.orc_demap, L.orc_hard_demap):
            fn.restype = None
            fn.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.c_uint64, C.POINTER(C.c_float)]
        L.orc_constellation.restype = C.POINTER(C.c_float)
        L.orc_constellation.argtypes = [C.c_uint32]

    def parity(self, masked: int, mode: int) -> int:
        return int(self.lib.orc_parity(C.c_uint64(masked), mode))

    def labels(self, code: Code) -> np.ndarray:
        out = np.zeros((code.S, 2), np.uint8)
        c = code._orc()
        self.lib.orc_labels(C.byref(c), _p(out, C.c_uint8))
        return out

    def decode_soft(self, code: Code, metrics: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(metrics, np.float32).reshape(-1, code.T, code.M)
        out = np.zeros((m.shape[0], code.out_bytes), np.uint8)
        c = code._orc()
        rc = self.lib.orc_decode_soft(C.byref(c), _p(m, C.c_float), m.shape[0], _p(out, C.c_uint8))
        assert rc == 0
        return out

    def decode_hard(self, code: Code, symbols: np.ndarray):
        s = np.ascontiguousarray(symbols, np.uint8).reshape(-1, code.T)
        out = np.zeros((s.shape[0], code.out_bytes), np.uint8)
        pm = np.zeros(s.shape[0], np.int32)
        c = code._orc()
        rc = self.lib.orc_decode_hard(C.byref(c), _p(s, C.c_uint8), s.shape[0], _p(out, C.c_uint8), _p(pm, C.c_int32))
        assert rc == 0
        return out, pm

    def encode(self, code: Code, payload: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(payload, np.uint8).reshape(-1, code.payload_bytes)
        sym = np.zeros((p.shape[0], code.T), np.uint8)
        c = code._orc()
        self.lib.orc_encode(C.byref(c), _p(p, C.c_uint8), p.shape[0], _p(sym, C.c_uint8))
        return sym

    def constellation(self, n: int) -> np.ndarray:
        ptr = self.lib.orc_constellation(n)
        return np.ctypeslib.as_array(ptr, shape=(1 << n, 2)).copy()

    def demap(self, n: int, iq: np.ndarray, hard: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(iq, np.float32).reshape(-1, 2)
        out = np.zeros((x.shape[0], 1 << n), np.float32)
        fn = self.lib.orc_hard_demap if hard else self.lib.orc_demap
        fn(n, _p(x, C.c_float), x.shape[0], _p(out, C.c_float))
        return out


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(_REF, f)) for f in ("libref_awgn.so", "libref_bsc.so"))


class Reference:
    """The UNMODIFIED reference objects behind oracle/ref_harness.c (K<=8, T<=255 only)."""

    def __init__(self):
        if not reference_available():
            raise RuntimeError("oracle/_ref/libref_*.so missing: run `make -C oracle` where /root/reference exists")
        self.awgn = C.CDLL(os.path.join(_REF, "libref_awgn.so"))
        self.awgn_hard = C.CDLL(os.path.join(_REF, "libref_awgn_hard.so"))
        self.bsc = C.CDLL(os.path.join(_REF, "libref_bsc.so"))
        for lib in (self.awgn, self.awgn_hard):
            lib.refh_awgn_run.restype = C.c_int
            lib.refh_awgn_run.argtypes = [C.POINTER(_RefhCode), C.c_uint64, C.c_uint32, C.c_float,
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
            lib.refh_awgn_decode.restype = C.c_int
            lib.refh_awgn_decode.argtypes = [C.POINTER(_RefhCode), C.POINTER(C.c_float), C.c_uint64, C.POINTER(C.c_uint8)]
            lib.refh_demap.restype = C.c_int
            lib.refh_demap.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.c_uint64, C.POINTER(C.c_float)]
        self.bsc.refh_bsc_run.restype = C.c_int
        self.bsc.refh_bsc_run.argtypes = [C.POINTER(_RefhCode), C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_uint8),
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        self.bsc.refh_bsc_decode.restype = C.c_int
        self.bsc.refh_bsc_decode.argtypes = [C.POINTER(_RefhCode), C.POINTER(C.c_uint8), C.c_uint64,
                                             C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        for lib in (self.awgn, self.awgn_hard, self.bsc):
            lib.refh_supported.restype = C.c_int
            lib.refh_supported.argtypes = [C.POINTER(_RefhCode)]
            lib.refh_get_code.restype = None
            lib.refh_get_code.argtypes = [C.c_uint32, C.POINTER(_RefhCode)]
            lib.refh_encode.restype = C.c_int
            lib.refh_encode.argtypes = [C.POINTER(_RefhCode), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8)]

    def supported(self, code: Code) -> bool:
        c = code._refh()
        return bool(self.awgn.refh_supported(C.byref(c)))

    def codebook(self, index: int) -> Code:
        c = _RefhCode()
        self.awgn.refh_get_code(index, C.byref(c))
        return Code(c.symlen_out, c.constr_len, c.block_len, [c.polynomials[i] for i in range(c.symlen_out)], PARITY_REF)

    def awgn_run(self, code: Code, nframes: int, seed: int, scaling: float, hard_demapper: bool = False):
        lib = self.awgn_hard if hard_demapper else self.awgn
        payload = np.zeros((nframes, code.payload_bytes), np.uint8)
        metrics = np.zeros((nframes, code.T, code.M), np.float32)
        out = np.zeros((nframes, code.out_bytes), np.uint8)
        c = code._refh()
        rc = lib.refh_awgn_run(C.byref(c), nframes, seed, scaling, _p(payload, C.c_uint8), _p(metrics, C.c_float), _p(out, C.c_uint8))
        assert rc == 0, rc
        return payload, metrics, out

    def awgn_decode(self, code: Code, metrics: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(metrics, np.float32).reshape(-1, code.T, code.M)
        out = np.zeros((m.shape[0], code.out_bytes), np.uint8)
        c = code._refh()
        rc = self.awgn.refh_awgn_decode(C.byref(c), _p(m, C.c_float), m.shape[0], _p(out, C.c_uint8))
        assert rc == 0, rc
        return out

    def demap(self, n: int, iq: np.ndarray, hard: bool = False) -> np.ndarray:
        lib = self.awgn_hard if hard else self.awgn
        x = np.ascontiguousarray(iq, np.float32).reshape(-1, 2)
        out = np.zeros((x.shape[0], 1 << n), np.float32)
        assert lib.refh_demap(n, _p(x, C.c_float), x.shape[0], _p(out, C.c_float)) == 0
        return out

    def bsc_run(self, code: Code, nframes: int, seed: int, crossover_ppm: int):
        payload = np.zeros((nframes, code.payload_bytes), np.uint8)
        sym = np.zeros((nframes, code.T), np.uint8)
        out = np.zeros((nframes, code.out_bytes), np.uint8)
        pm = np.zeros(nframes, np.int32)
        c = code._refh()
        rc = self.bsc.refh_bsc_run(C.byref(c), nframes, seed, crossover_ppm, _p(payload, C.c_uint8), _p(sym, C.c_uint8),
                                   _p(out, C.c_uint8), _p(pm, C.c_int32))
        assert rc == 0, rc
        return payload, sym, out, pm

    def bsc_decode(self, code: Code, symbols: np.ndarray):
        s = np.ascontiguousarray(symbols, np.uint8).reshape(-1, code.T)
        out = np.zeros((s.shape[0], code.out_bytes), np.uint8)
        pm = np.zeros(s.shape[0], np.int32)
        c = code._refh()
        rc = self.bsc.refh_bsc_decode(C.byref(c), _p(s, C.c_uint8), s.shape[0], _p(out, C.c_uint8), _p(pm, C.c_int32))
        assert rc == 0, rc
        return out, pm

    def encode(self, code: Code, payload: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(payload, np.uint8).reshape(-1, code.payload_bytes)
        sym = np.zeros((p.shape[0], code.T), np.uint8)
        c = code._refh()
        assert self.bsc.refh_encode(C.byref(c), _p(p, C.c_uint8), p.shape[0], _p(sym, C.c_uint8)) == 0
        return sym
