#!/usr/bin/env python
"""Per-kernel register / spill table from a ptxas -v log:  python tools/regs_report.py build/obj/vdec_regs.cu.log"""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
names = re.findall(r"Compiling entry function '(\S+)'", txt)
spills = re.findall(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", txt)
regs = re.findall(r"Used (\d+) registers", txt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
for n, s, r in zip(dem, spills, regs):
    n = re.sub(r"\(unsigned long\)|\(int\)|\(bool\)|vd::", "", n.split("(const")[0])
    print(f"{r:>4} regs  spill st/ld {s[1]:>4}/{s[2]:<4} {n[:150]}")
