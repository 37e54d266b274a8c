#!/usr/bin/env python
"""Build libvdec.so with extra -D flags into build/<name>/ for A/B runs (VDEC_LIB_DIR=build/<name>).
    python tools/build_variant.py mb5 -DVDEC_MINB_REF=5 -DVDEC_TRACE_CHUNK=8
    python tools/build_variant.py xg3 --only=vdec_regs.cu -DVDEC_RING_XG_M8=3     (other objects come from build/obj)"""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "convolutional-codes_b200", "csrc")
name, extra = sys.argv[1], sys.argv[2:]
only = [a.split("=", 1)[1].split(",") for a in extra if a.startswith("--only=")]
only = only[0] if only else None
extra = [a for a in extra if not a.startswith("--only=")]
out = os.path.join(ROOT, "build", name)
os.makedirs(out, exist_ok=True)
flags = ["-std=c++20", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "-I", os.path.join(ROOT, "build", "obj"), *extra]          # build/obj holds the generated vdec_jit_sources.inc


def comp(n):
    if only is not None and n not in only:
        return os.path.join(ROOT, "build", "obj", n.replace(".cu", ".o"))
    o = os.path.join(out, n.replace(".cu", ".o"))
    r = subprocess.run(["nvcc", *flags, "-c", os.path.join(SRC, n), "-o", o], capture_output=True, text=True)
    open(o + ".log", "w").write(r.stderr)
    assert r.returncode == 0, r.stderr[-3000:]
    return o


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, ["vdec_api.cu", "vdec_regs.cu", "vdec_regs2.cu", "vdec_regsw.cu", "vdec_tile.cu", "vdec_tile2.cu", "vdec_generic.cu", "vdec_synth.cu", "vdec_sweep.cu",
                              "vdec_jit.cu"]))
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(out, "libvdec.so"), *objs,
                "-Xlinker", "-Bsymbolic", "-ldl"], check=True)
print("built", out)
