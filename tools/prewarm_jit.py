#!/usr/bin/env python
"""Compile (NVRTC, no GPU needed) the run-time instantiated kernels the test-suite and smoke() ask for into the
opt-in on-disk cache $VDEC_JIT_CACHE (default here: build/jit_cache, which travels to the GPU box with the repo
snapshot like the built .so files), so a GPU run spends its time decoding instead of compiling.
    python tools/prewarm_jit.py [-j 8]"""
import argparse
import os
import shutil
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CACHE = os.path.join(ROOT, "build", "jit_cache")


def jobs():
    """(n, K, L, octal polynomials, parity, metric, flags, metric_max, iq)"""
    out = []
    src = open(os.path.join(ROOT, "tests", "test_gpu_parity.py")).read()
    ns = {}
    start = src.index("JIT_CASES = [")
    exec(src[start:src.index("]\n", start) + 2], ns)
    SOFT, HARD, QUANT, P32 = 0, 1, 2, 4
    for n, K, L, octal, _ in ns["JIT_CASES"]:
        for parity in (0, 1):
            out.append((n, K, L, octal, parity, SOFT, 0, 0, False))
            out.append((n, K, L, octal, parity, HARD, 0, 0, False))
            if K <= 7:
                out.append((n, K, L, octal, parity, HARD, P32, 0, False))
            if n == 2:
                out.append((n, K, L, octal, parity, QUANT, 0, 7, False))
    out.append((2, 5, 64, ("23", "35"), 1, SOFT, 0, 0, False))          # __graft_entry__.smoke()
    for parity in (0, 1):                                                # fuzz / raw-I/Q cases
        out.append((2, 6, 77, ("53", "75"), parity, SOFT, 0, 0, False))
        out.append((2, 6, 77, ("53", "75"), parity, SOFT, 0, 0, True))
    for K, octal, L in ((8, ("225", "331", "367"), 100), (10, ("1117", "1365", "1633"), 130), (13, ("10533", "17661", "11271"), 40)):
        out.append((3, K, L, octal, 1, SOFT, 0, 0, False))               # rate 1/3 register-tile kernels
    return out


def run(job):
    os.environ.setdefault("VDEC_JIT_CACHE", CACHE)
    os.environ["VDEC_QUIET"] = "1"
    import vdec_pkg
    cc = vdec_pkg.load()
    n, K, L, octal, parity, metric, flags, mmax, iq = job
    code = cc.CodeParam(n, K, L, tuple(cc.codes.left_align(o, K) for o in octal))
    try:
        return job, cc.jit_check(code, metric, parity, flags, mmax, iq=iq)
    except Exception as e:          # a failing compile is the tests' business; the cache just stays cold for it
        return job, repr(e)


def _sources_digest() -> str:
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "convolutional-codes_b200", "csrc")
    for n in sorted(os.listdir(d)):
        if n.endswith((".cuh", ".h")):
            h.update(open(os.path.join(d, n), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "vdec.h"), "rb").read())
    h.update(open(os.path.join(ROOT, "convolutional-codes_b200", "_build.py"), "rb").read())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--keep", action="store_true", help="do not drop cache entries of other source versions")
    args = ap.parse_args()
    # the cache key hashes the kernel headers: entries of earlier source versions are dead weight in the GPU-box
    # snapshot, so the directory is rebuilt from scratch unless every job is already served from it (--keep)
    if not args.keep and os.path.isdir(CACHE):
        stamp = os.path.join(CACHE, ".sources")
        cur = _sources_digest()
        if not os.path.exists(stamp) or open(stamp).read() != cur:
            shutil.rmtree(CACHE)
    os.makedirs(CACHE, exist_ok=True)
    with open(os.path.join(CACHE, ".sources"), "w") as f:
        f.write(_sources_digest())
    with ProcessPoolExecutor(args.j) as ex:
        res = list(ex.map(run, jobs()))
    bad = [r for r in res if not isinstance(r[1], int)]
    print(f"prewarmed {len(res) - len(bad)} kernels into {CACHE}" + (f", {len(bad)} failed: {bad[:2]}" if bad else ""))


if __name__ == "__main__":
    main()
