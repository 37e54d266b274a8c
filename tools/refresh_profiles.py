#!/usr/bin/env python
"""Copy the evidence of one tools/gpu_session.sh call from gpurun_out/ into profiles/ and refresh profiles/ncu_traffic.json.
    python tools/refresh_profiles.py r2y r02y [old profile tag to remove, e.g. r02x]"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, ptag = sys.argv[1], sys.argv[2]
old = sys.argv[3] if len(sys.argv) > 3 else None
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
NAMES = {"bench.json": "bench_1gpu.json", "bench_ref.json": "bench_reference_arm.json", "launches.csv": "launches.csv"}
NCU = {"full_ref_iq": ("vit_regs_ref_and_iq", "vit_regs_kernel K7 (171,133) ref labels, metrics input (1st block) and raw-I/Q feeder "
                       "(2nd block), 606208 frames each"),
       "full_lin_r13": ("vit_regs_lin_and_r13", "vit_regs_kernel K7 (171,133) true-parity labels (606208 frames) and K7 r=1/3 ref labels "
                        "(303104 frames)")}
if old:
    for suffix in list(NAMES.values()) + [v[0] + "_ncu_full_summary.txt" for v in NCU.values()]:
        p = os.path.join(P, f"{old}_{suffix}")
        if os.path.exists(p):
            os.remove(p)
for src, dst in NAMES.items():
    shutil.copy(os.path.join(G, f"{tag}_{src}"), os.path.join(P, f"{ptag}_{dst}"))
for rep, (name, what) in NCU.items():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join("gpurun_out", f"{tag}_{rep}.ncu-rep"),
                          f"{tag}, final kernels: {what}; ncu --set full --clock-control none"], capture_output=True, text=True, check=True,
                         cwd=ROOT).stdout
    open(os.path.join(P, f"{ptag}_{name}_ncu_full_summary.txt"), "w").write(out)
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"{tag}_full_ref_iq.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True,
                     check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, first = rows[0], rows[1], rows[2]
def val(k):
    i = hdr.index(k)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
    return float(first[i].replace(",", "")) * scale
tp = os.path.join(P, "ncu_traffic.json")
d = json.load(open(tp))
d["source"] = (f"profiles/{ptag}_vit_regs_ref_and_iq_ncu_full_summary.txt (ncu --set full --clock-control none, one launch of "
               f"{d['frames']} frames, first block)")
d["dram__bytes_read.sum"], d["dram__bytes_write.sum"] = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
d["dram_bytes_per_frame"] = (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / d["frames"]
json.dump(d, open(tp, "w"), indent=1)
print("dram bytes per frame:", round(d["dram_bytes_per_frame"], 1))
