#!/usr/bin/env python
"""Instruction histogram of the hottest loop (the longest backward branch) of one kernel in an object file.
    python tools/sass_loop.py build/obj/vdec_regs.o 'CodeDefILi7ELi2E.*ELb1EEENS_7SoftF32'"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], re.compile(sys.argv[2])
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0]
    if not pat.search(name):
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_]+)([.\w]*)\s*([^;]*);", f)
    best = (0, 0, 0)
    for addr, pred, op, mod, args in ins:
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", args)
            if m and int(m.group(1), 16) < int(addr, 16):
                span = int(addr, 16) - int(m.group(1), 16)
                if span > best[0]:
                    best = (span, int(m.group(1), 16), int(addr, 16))
    span, lo, hi = best
    hist = collections.Counter()
    for addr, pred, op, mod, args in ins:
        a = int(addr, 16)
        if lo <= a <= hi:
            hist[("@p " if pred.strip() and "PT" not in pred else "") + op] += 1
    n = sum(hist.values())
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(unsigned long\)|\(int\)|\(bool\)|vd::", "", dem.split("(const")[0].split("(void const")[0])
    print(f"{dem[:200]}\n  loop {lo:#x}..{hi:#x}: {n} instructions")
    print("  " + ", ".join(f"{k} {v}" for k, v in hist.most_common(40)))
