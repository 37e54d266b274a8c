#!/usr/bin/env python
"""Small decode of every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Sizes are tiny (the tools slow kernels down 10-100x); results are still compared with the CPU checker."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import vdec_pkg
    import oracle
    cc = vdec_pkg.load()
    orc = oracle.Oracle()
    C = cc.codes
    soft = [(C.K7_R12.with_block_len(70), cc.PARITY_REF, 67, 0), (C.K7_R13.with_block_len(33), cc.PARITY_LINEAR, 40, 0),
            (cc.get_code(0), cc.PARITY_REF, 100, 0), (C.K8_R12.with_block_len(40), cc.PARITY_LINEAR, 19, 0),
            (C.K9_R12.with_block_len(50), cc.PARITY_LINEAR, 21, 0), (C.K9_R12.with_block_len(50), cc.PARITY_LINEAR, 21, cc.PREFER_LANEGROUP),
            (C.K12_R12.with_block_len(31), cc.PARITY_LINEAR, 9, 0), (C.K13_R12.with_block_len(20), cc.PARITY_LINEAR, 5, 0),
            (C.K14_R12.with_block_len(14), cc.PARITY_LINEAR, 3, 0), (C.K15_R12.with_block_len(13), cc.PARITY_LINEAR, 2, 0),
            (C.K9_R12.with_block_len(30), cc.PARITY_REF, 4, 0)]
    for code, parity, F, flags in soft:
        dec = cc.BatchDecoder(code, cc.SOFT_F32, parity, flags=flags)
        _, met = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(3.0))
        out = dec.decode(met).cpu().numpy()
        oc = oracle.Code(code.n, code.K, code.L, code.polynomials, parity)
        assert np.array_equal(out, orc.decode_soft(oc, met.cpu().numpy())), dec.kernel_name
        print("ok", dec.kernel_name, flush=True)
        dec.close()
    for code, parity, F, flags in [(C.K7_R12.with_block_len(90), cc.PARITY_LINEAR, 65, 0), (cc.get_code(0), cc.PARITY_REF, 130, 0),
                                   (C.K7_R12.with_block_len(90), cc.PARITY_LINEAR, 33, cc.PREFER_32BIT)]:
        dec = cc.BatchDecoder(code, cc.HARD_U16, parity, flags=flags)
        _, sym = dec.synth_bsc(F, 0, 1, 30000)
        out, pm = dec.decode(sym)
        oc = oracle.Code(code.n, code.K, code.L, code.polynomials, parity)
        e_out, e_pm = orc.decode_hard(oc, sym.cpu().numpy())
        assert np.array_equal(out.cpu().numpy(), e_out) and np.array_equal(pm.cpu().numpy(), e_pm), dec.kernel_name
        print("ok", dec.kernel_name, flush=True)
        dec.close()
    code = C.K7_R12.with_block_len(64)
    dec = cc.BatchDecoder(code, cc.SOFT_F32, cc.PARITY_LINEAR)
    _, x = dec.synth_awgn(40, 0, 1, cc.awgn_sigma(4.0))
    dec.decode_stream(x.reshape(-1, code.M), 256, 64, start_known=True)
    torch.cuda.synchronize()
    print("ok stream", flush=True)
    # raw-I/Q entry (demapper in the feeder), I/Q-only source, hard demapper, the C sweep entry
    for code, parity, F in ((C.K7_R12.with_block_len(61), cc.PARITY_REF, 45), (C.K7_R13.with_block_len(35), cc.PARITY_LINEAR, 33),
                            (cc.get_code(5), cc.PARITY_REF, 70)):
        d = cc.BatchDecoder(code, cc.SOFT_F32, parity)
        _, iq = d.synth_awgn(F, 3, 1, cc.awgn_sigma(3.0), want_iq=True, want_metrics=False)
        oc = oracle.Code(code.n, code.K, code.L, code.polynomials, parity)
        met = orc.demap(code.n, iq.cpu().numpy().reshape(-1, 2)).reshape(F, code.T, code.M)
        assert np.array_equal(d.decode_iq(iq).cpu().numpy(), orc.decode_soft(oc, met)), d.iq_kernel_name
        hm = cc.demap(code.n, iq, hard=True)
        assert np.array_equal(hm.cpu().numpy().view(np.uint32), orc.demap(code.n, iq.cpu().numpy().reshape(-1, 2), hard=True).reshape(hm.shape).view(np.uint32))
        print("ok", d.iq_kernel_name, flush=True)
        d.close()
    got = cc.sweep(cc.get_code(3), [3.0], 600, cc.CHANNEL_AWGN, cc.PARITY_REF, seed=2, ndev=1, chunk_frames=256)
    assert int(got[0][0]) == 600 * 40
    print("ok sweep", got.tolist(), flush=True)


if __name__ == "__main__":
    main()
