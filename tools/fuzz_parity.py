#!/usr/bin/env python
"""Randomised differential test on the GPU: random codes (K, rate, polynomials, label mode), frame lengths, frame counts
and channel noise through the C ABI against the CPU oracle, for a bounded time.  Every output bit (tail included) and
every path metric must be equal.  Polynomials are random, so almost every case exercises a kernel compiled at run time.

    python tools/fuzz_parity.py --seconds 300 --seed 1 > results/fuzz_parity.json
"""
import argparse
import collections
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import vdec_pkg
    import oracle
    from conftest import make_hard_inputs, make_soft_inputs, ocode
    cc = vdec_pkg.load()
    orc = oracle.Oracle()
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    stats = collections.Counter()
    families = collections.Counter()
    failures = []
    while time.time() < t_end:
        K = int(rng.integers(3, 13))
        n = int(rng.choice([1, 2, 2, 2, 3, 3]))
        L = int(rng.choice([1, 2, 7, 40, 97, 249, 300, 1024])) if rng.random() < 0.5 else int(rng.integers(1, 400))
        if K >= 11:
            L = min(L, 200)
        F = int(rng.integers(1, 150 if K <= 9 else 12))
        # left-aligned polynomials: tap on the newest bit always, on the oldest bit for at least the first one
        polys = []
        for j in range(n):
            v = int(rng.integers(0, 1 << K)) | (1 << (K - 1)) | (1 if j == 0 else 0)
            polys.append((v << (64 - K)) & 0xFFFFFFFFFFFFFFFF)
        parity = int(rng.integers(0, 2))
        metric = int(rng.choice([cc.SOFT_F32, cc.SOFT_F32, cc.HARD_U16, cc.SOFT_U8]))
        flags = int(rng.choice([0, 0, 0, cc.NO_JIT, cc.PREFER_32BIT]))
        code = cc.CodeParam(n, K, L, tuple(polys))
        oc = ocode(code, parity)
        seed = int(rng.integers(0, 1 << 31))
        case = {"K": K, "n": n, "L": L, "F": F, "polys": [oct(p >> (64 - K)) for p in polys], "parity": parity, "metric": metric, "flags": flags}
        try:
            if metric == cc.HARD_U16:
                _, sym = make_hard_inputs(orc, oc, F, float(rng.choice([0.0, 0.02, 0.08, 0.3])), seed)
                exp, exp_pm = orc.decode_hard(oc, sym)
                dec = cc.BatchDecoder(code, metric, parity, flags=flags)
                o, p = dec.decode(torch.from_numpy(sym).cuda())
                ok = np.array_equal(o.cpu().numpy(), exp) and np.array_equal(p.cpu().numpy(), exp_pm)
            else:
                _, met = make_soft_inputs(orc, oc, F, float(rng.choice([0.2, 0.45, 0.7, 1.5])), seed, hard_demap=bool(rng.random() < 0.2))
                if metric == cc.SOFT_U8:
                    qmax = int(rng.choice([2, 7, 31]))
                    q = np.minimum(np.rint(met * 2.0), qmax).astype(np.uint8)
                    exp = orc.decode_soft(oc, q.astype(np.float32))
                    dec = cc.BatchDecoder(code, metric, parity, flags=flags, metric_max=qmax)
                    o = dec.decode(torch.from_numpy(q).cuda())
                elif rng.random() < 0.35:
                    # raw-I/Q entry: demapper inside the decoder kernel (register kernels) or demap kernel + decoder
                    cst = orc.constellation(n)
                    r2 = np.random.default_rng(seed)
                    pl = r2.integers(0, 256, size=(F, oc.payload_bytes), dtype=np.uint8)
                    iq = cst[orc.encode(oc, pl)].astype(np.float32) + (float(rng.choice([0.0, 0.3, 0.8])) * r2.standard_normal((F, oc.T, 2))).astype(np.float32)
                    exp = orc.decode_soft(oc, orc.demap(n, iq.reshape(-1, 2)).reshape(F, oc.T, oc.M))
                    dec = cc.BatchDecoder(code, metric, parity, flags=flags)
                    o = dec.decode_iq(torch.from_numpy(np.ascontiguousarray(iq)).cuda())
                    case["entry"] = "iq: " + dec.iq_kernel_name.split("<")[0]
                    stats["iq_cases"] += 1
                else:
                    exp = orc.decode_soft(oc, met)
                    dec = cc.BatchDecoder(code, metric, parity, flags=flags)
                    o = dec.decode(torch.from_numpy(np.ascontiguousarray(met)).cuda())
                ok = np.array_equal(o.cpu().numpy(), exp)
            name = dec.kernel_name
            dec.close()
        except Exception as e:                       # a config the library refuses is a finding too
            ok, name = False, "exception: " + repr(e)[:200]
        fam = name.split("<")[0] + (" (nvrtc)" if name.endswith("nvrtc>") else "")
        families[fam] += 1
        stats["cases"] += 1
        stats["frames"] += F
        stats["decoded_bits_compared"] += F * (L + K - 1)
        if not ok:
            stats["mismatching_cases"] += 1
            failures.append({**case, "kernel": name})
    print(json.dumps({"seed": args.seed, "seconds": args.seconds, **stats, "mismatching_cases": stats["mismatching_cases"],
                      "kernel_families": dict(families), "failures": failures[:20]}, indent=1))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
