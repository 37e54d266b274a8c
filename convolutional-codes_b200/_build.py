"""Builds the in-tree native libraries (sm_100a only).

    lib/libvdec.so        CUDA kernels + the batched C ABI (include/vdec.h)
    lib/libvdec_awgn.so   legacy soft-decision plugin ABI (include/vdec_legacy_awgn.h), C
    lib/libvdec_bsc.so    legacy hard-decision plugin ABI (include/vdec_legacy_bsc.h), C

nvcc cross-compiles without a GPU; the .so files travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(os.path.dirname(HERE), "build", "obj")

CU_SOURCES = ["vdec_api.cu", "vdec_regs.cu", "vdec_regs2.cu", "vdec_regsw.cu", "vdec_tile.cu", "vdec_tile2.cu", "vdec_generic.cu", "vdec_synth.cu"]
NVCC_FLAGS = ["-std=c++20", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _sources():
    inc = os.path.join(os.path.dirname(HERE), "include")
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(inc, f) for f in os.listdir(inc)]
    return [p for p in out if os.path.isfile(p)]


def is_fresh() -> bool:
    stamp = os.path.join(LIB, ".stamp")
    libs = [os.path.join(LIB, n) for n in ("libvdec.so", "libvdec_awgn.so", "libvdec_bsc.so")]
    if not all(os.path.exists(p) for p in libs) or not os.path.exists(stamp):
        return False
    return open(stamp).read().strip() == _digest(_sources())


def build(force: bool = False, verbose: bool = False) -> None:
    if not force and is_fresh():
        return
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_cu(name: str) -> str:
        obj = os.path.join(OBJ, name.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, name), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, name + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {name}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(CU_SOURCES)) as ex:
        objs = list(ex.map(compile_cu, CU_SOURCES))
    libvdec = os.path.join(LIB, "libvdec.so")
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", libvdec, *objs,
                    "-Xlinker", "-Bsymbolic"], check=True)
    cc = os.environ.get("CC") or shutil.which("gcc") or "gcc"
    for shim in ("awgn", "bsc"):
        subprocess.run([cc, "-O2", "-fPIC", "-Wall", "-Wextra", "-shared", "-Wl,-Bsymbolic",
                        "-o", os.path.join(LIB, f"libvdec_{shim}.so"), os.path.join(CSRC, f"shim_{shim}.c"),
                        "-L", LIB, "-lvdec", "-Wl,-rpath,$ORIGIN"], check=True)
    with open(os.path.join(LIB, ".stamp"), "w") as f:
        f.write(_digest(_sources()))


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built:", sorted(os.listdir(LIB)))
