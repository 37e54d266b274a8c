#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config shape (not the headline bench):
decoded Gbit/s and ACS Gstate-updates/s per kernel family, CUDA-event timed, inputs in HBM.

    python tools/bench_configs.py [--quick]
"""
import argparse
import re
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import vdec_pkg
    cc = vdec_pkg.load()
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="filter on the case name: substring, or regular expression")
    args = ap.parse_args()
    q = 8 if args.quick else 1
    C = cc.codes
    cases = [  # name, code, metric, parity, frames, flags
        ("C1 hard K3 cb0 L40", cc.get_code(0), cc.HARD_U16, cc.PARITY_REF, (1 << 20) // q, 0),
        ("C1' soft K3 cb0 L40", cc.get_code(0), cc.SOFT_F32, cc.PARITY_REF, (1 << 20) // q, 0),
        ("cb3 soft K6 L40", cc.get_code(3), cc.SOFT_F32, cc.PARITY_REF, (1 << 20) // q, 0),
        ("C2 soft K7 r1/2 L1024 ref", C.K7_R12, cc.SOFT_F32, cc.PARITY_REF, 1000000 // q, 0),
        ("C2 soft K7 r1/2 L1024 linear", C.K7_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 1000000 // q, 0),
        ("C2h hard K7 r1/2 L1024", C.K7_R12, cc.HARD_U16, cc.PARITY_LINEAR, 1000000 // q, 0),
        ("C2h hard K7 r1/2 L1024 ref-labels", C.K7_R12, cc.HARD_U16, cc.PARITY_REF, 1000000 // q, 0),
        ("f3 quant K7 r1/2 L1024 ref-labels (u8 metrics <= 7)", C.K7_R12, cc.SOFT_U8, cc.PARITY_REF, 1000000 // q, 0),
        ("C2h hard K7 r1/2 L1024 (32-bit kernel)", C.K7_R12, cc.HARD_U16, cc.PARITY_LINEAR, 1000000 // q, cc.PREFER_32BIT),
        ("f3 quant K7 r1/2 L1024 (u8 metrics <= 7)", C.K7_R12, cc.SOFT_U8, cc.PARITY_LINEAR, 1000000 // q, 0),
        ("C3 soft K7 r1/3 L1024 ref", C.K7_R13, cc.SOFT_F32, cc.PARITY_REF, 511488 // q, 0),
        ("C3 soft K7 r1/3 L1024 linear", C.K7_R13, cc.SOFT_F32, cc.PARITY_LINEAR, 511488 // q, 0),
        ("C3h hard K7 r1/3 L1024 ref-labels", C.K7_R13, cc.HARD_U16, cc.PARITY_REF, 1000000 // q, 0),
        ("C3h hard K7 r1/3 L1024", C.K7_R13, cc.HARD_U16, cc.PARITY_LINEAR, 1000000 // q, 0),
        # polynomials outside the registries: the same kernel families, compiled at run time by NVRTC (vdec_jit.cu)
        ("K7 (165,127) soft L1024, nvrtc", cc.CodeParam(2, 7, 1024, (C.left_align("165", 7), C.left_align("127", 7))), cc.SOFT_F32, cc.PARITY_LINEAR, 1000000 // q, 0),
        ("K7 (165,127) soft L1024, table-driven", cc.CodeParam(2, 7, 1024, (C.left_align("165", 7), C.left_align("127", 7))), cc.SOFT_F32, cc.PARITY_LINEAR, 8192 // q, cc.NO_JIT),
        ("K7 (165,127) hard L1024, nvrtc", cc.CodeParam(2, 7, 1024, (C.left_align("165", 7), C.left_align("127", 7))), cc.HARD_U16, cc.PARITY_LINEAR, 1000000 // q, 0),
        ("K9 (657,435) soft L1024, nvrtc", cc.CodeParam(2, 9, 1024, (C.left_align("657", 9), C.left_align("435", 9))), cc.SOFT_F32, cc.PARITY_LINEAR, 131072 // q, 0),
        ("C2 generic K7 (table-driven)", C.K7_R12, cc.SOFT_F32, cc.PARITY_REF, 8192 // q, cc.FORCE_GENERIC),
        ("K8 soft L1024", C.K8_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 262144 // q, 0),
        ("K8 soft L1024 (lane-group)", C.K8_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 262144 // q, cc.PREFER_LANEGROUP),
        ("C4 soft K9 L1024", C.K9_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 131072 // q, 0),
        ("C4 soft K9 L1024 ref-labels", C.K9_R12, cc.SOFT_F32, cc.PARITY_REF, 131072 // q, 0),
        ("C4 soft K9 L1024 ref-labels (table-driven)", C.K9_R12, cc.SOFT_F32, cc.PARITY_REF, 16384 // q, cc.FORCE_GENERIC),
        ("K9 hard L1024", C.K9_R12, cc.HARD_U16, cc.PARITY_LINEAR, 262144 // q, 0),
        ("K9 hard L1024 (table-driven)", C.K9_R12, cc.HARD_U16, cc.PARITY_LINEAR, 16384 // q, cc.FORCE_GENERIC),
        ("K9 quant L1024 (u8 metrics <= 7)", C.K9_R12, cc.SOFT_U8, cc.PARITY_LINEAR, 262144 // q, 0),
        ("K15 hard L1024", C.K15_R12, cc.HARD_U16, cc.PARITY_LINEAR, 4736 // q, 0),
        ("C4 soft K11 L1024", C.K11_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 32768 // q, 0),
        ("C4 soft K9 L1024 (lane-group)", C.K9_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 131072 // q, cc.PREFER_LANEGROUP),
        ("C4 soft K11 L1024 (lane-group)", C.K11_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 32768 // q, cc.PREFER_LANEGROUP),
        ("K10 soft L1024", C.K10_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 65536 // q, 0),
        ("K12 soft L1024", C.K12_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 18944 // q, 0),
        ("K14 soft L1024", C.K14_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 4736 // q, 0),
        ("C4 soft K13 L1024", C.K13_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 9472 // q, 0),
        ("C4 soft K15 L1024", C.K15_R12, cc.SOFT_F32, cc.PARITY_LINEAR, 2368 // q, 0),
        ("C4 soft K15 L1024 ref-labels", C.K15_R12, cc.SOFT_F32, cc.PARITY_REF, 2368 // q, 0),
        ("C4 soft K15 L1024 ref-labels (table-driven)", C.K15_R12, cc.SOFT_F32, cc.PARITY_REF, 1024 // q, cc.FORCE_GENERIC),
    ]
    if not args.only or args.only in "stream K7 sliding-window":
        # one unterminated stream of 2^28 steps, windows of 1024 payload + 2x128 overlap steps
        code = C.K7_R12
        nsteps = (1 << 28) // q
        dec = cc.BatchDecoder(code, cc.SOFT_F32, cc.PARITY_LINEAR)
        nb = nsteps // code.T
        _, x = dec.synth_awgn(nb, 0, 1, cc.awgn_sigma(4.0))       # frames back to back = a long symbol stream
        x = x.reshape(nb * code.T, code.M)
        for blk, dep in ((1024, 128), (4096, 128)):
            dec.decode_stream(x, blk, dep, start_known=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                dec.decode_stream(x, blk, dep, start_known=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(json.dumps({"case": f"stream K7 sliding-window block={blk} depth={dep}", "kernel": dec.kernel_name,
                              "steps": x.shape[0], "ms": round(ms, 3), "gbit_s": round(x.shape[0] / ms / 1e6, 3)}), flush=True)
        dec.close(); del x
        torch.cuda.empty_cache()
    for name, code, metric, parity, F, flags in cases:
        if args.only and args.only not in name and not re.search(args.only, name):
            continue
        dec = cc.BatchDecoder(code, metric, parity, flags=flags, metric_max=7 if metric == cc.SOFT_U8 else 0)
        if metric == cc.SOFT_F32:
            payload, x = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(4.0))
        elif metric == cc.SOFT_U8:
            payload, xf = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(4.0))
            x = torch.clamp(torch.round(xf * 2.0), max=7).to(torch.uint8)      # 3-bit quantised squared distances
            del xf
        else:
            payload, x = dec.synth_bsc(F, 0, 1, 20000)
        out = torch.empty((F, dec.out_bytes), dtype=torch.uint8, device="cuda")
        pm = torch.empty((F,), dtype=torch.int32, device="cuda")
        for _ in range(2):
            dec.decode(x, out=out, path_metric=pm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            dec.decode(x, out=out, path_metric=pm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cnt = dec.count_errors(out, payload).cpu().tolist()
        print(json.dumps({"case": name, "kernel": dec.kernel_name, "frames": F, "ms": round(ms, 3),
                          "gbit_s": round(F * code.L / ms / 1e6, 3),
                          "gstate_upd_s": round(F * code.T * code.S / ms / 1e6, 1),
                          "ber": cnt[1] / max(cnt[0], 1)}), flush=True)
        dec.close()
        del x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
