"""TEST INFRASTRUCTURE — regenerates tests/golden/*.npz from the UNMODIFIED reference.

Run in a container that has /root/reference:   python oracle/gen_golden.py
Each file holds what the reference's decoder object consumed and produced for a fixed srand()
seed (oracle/ref_harness.c wires the objects as the reference drivers do):
    payload  uint8 [F][(L+7)/8]   the driver's random payload (main.c:217-219)
    inputs   float32 [F][T][M]    demapper output fed to decoder_input   (soft)
             uint8  [F][T]        bit-flipped encoder output             (hard)
    outputs  uint8 [F][(T+7)/8]   the buffer handed to the output callback, all T bits
    path_metric int32 [F]         (hard only)
    meta     n, K, L, polynomials, seed, channel parameter
The reference repository ships no test vectors of its own, so these are the pin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SIGMA_4DB = np.float32(0.446154216921401)   # AWGN-channel/main.c:158
SIGMA_2DB = np.float32(0.561674881261479)

K7 = [0xF200000000000000, 0xB600000000000000]
K7R13 = [0xB600000000000000, 0xF200000000000000, 0xEA00000000000000]
K8 = [0xF900000000000000, 0xA700000000000000]   # (371,247)_8: largest K the reference can run


def main():
    R = oracle.Reference()
    os.makedirs(OUT, exist_ok=True)
    cases = []
    for idx in (0, 1, 2, 3, 5):
        cases.append((f"cb{idx}", R.codebook(idx), 24))
    cases.append(("k7r12_L249", oracle.Code(2, 7, 249, K7), 12))
    cases.append(("k7r13_L200", oracle.Code(3, 7, 200, K7R13), 8))
    cases.append(("k8r12_L100", oracle.Code(2, 8, 100, K8), 8))
    for name, code, F in cases:
        meta = dict(n=code.n, K=code.K, L=code.L, polynomials=np.array(code.polys, np.uint64))
        for tag, sigma, hard_demap in (("soft4dB", SIGMA_4DB, False), ("soft2dB", SIGMA_2DB, False), ("harddemap4dB", SIGMA_4DB, True)):
            seed = 1000 + len(name) * 7 + int(sigma * 1000)
            payload, metrics, out = R.awgn_run(code, F, seed, float(sigma), hard_demapper=hard_demap)
            np.savez_compressed(os.path.join(OUT, f"{name}_{tag}.npz"), kind="soft", payload=payload, inputs=metrics,
                                outputs=out, seed=seed, sigma=sigma, **meta)
        for ppm in (12500, 78125):
            seed = 77 + ppm
            payload, sym, out, pm = R.bsc_run(code, F, seed, ppm)
            np.savez_compressed(os.path.join(OUT, f"{name}_bsc{ppm}.npz"), kind="hard", payload=payload, inputs=sym,
                                outputs=out, path_metric=pm, seed=seed, crossover_ppm=ppm, **meta)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {total / 1024:.0f} KiB to {OUT}")


if __name__ == "__main__":
    main()
