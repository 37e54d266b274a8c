import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the library's on-disk kernel cache is opt-in: the suite opts in with an in-tree directory (tools/prewarm_jit.py fills
# it on the CPU at build time, so a GPU run does not spend minutes in NVRTC)
os.environ.setdefault("VDEC_JIT_CACHE", os.path.join(ROOT, "build", "jit_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def cc():
    """the product package (convolutional-codes_b200), built if necessary"""
    import vdec_pkg
    vdec_pkg.build_module().build()
    return vdec_pkg.load()


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.reference_available():
        pytest.skip("oracle/_ref/libref_*.so not built (needs /root/reference at build time)")
    return oracle.Reference()


def golden_files(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        if kind is None or str(np.load(p)["kind"]) == kind:
            out.append(p)
    return out


def load_golden(path):
    import oracle
    z = np.load(path)
    code = oracle.Code(int(z["n"]), int(z["K"]), int(z["L"]), [int(x) for x in z["polynomials"]], oracle.PARITY_REF)
    return z, code


def to_param(cc, code, L=None):
    """oracle.Code -> product CodeParam"""
    return cc.CodeParam(code.n, code.K, code.L if L is None else L, tuple(code.polys))


def ocode(code, parity=0, L=None):
    """product CodeParam -> oracle.Code"""
    import oracle
    return oracle.Code(code.n, code.K, code.L if L is None else L, code.polynomials, parity)


def make_soft_inputs(orc, code, F, sigma, seed, hard_demap=False):
    """random payload -> oracle encoder -> constellation -> AWGN -> oracle demapper (CPU test signal)"""
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, size=(F, code.payload_bytes), dtype=np.uint8)
    sym = orc.encode(code, payload)
    cst = orc.constellation(code.n)
    iq = cst[sym].astype(np.float32) + (sigma * rng.standard_normal((F, code.T, 2))).astype(np.float32)
    met = orc.demap(code.n, iq.reshape(-1, 2), hard=hard_demap).reshape(F, code.T, code.M)
    return payload, met


def make_hard_inputs(orc, code, F, p, seed):
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, size=(F, code.payload_bytes), dtype=np.uint8)
    sym = orc.encode(code, payload)
    flips = (rng.random((F, code.T, code.n)) < p)
    mask = (flips * (1 << np.arange(code.n))).sum(axis=2).astype(np.uint8)
    return payload, sym ^ mask
