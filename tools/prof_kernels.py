#!/usr/bin/env python
"""One launch of each kernel the bench line's numbers rest on, for `ncu --set full` (run under ncu, never timed):
    ncu --set full --clock-control none --import-source on -k regex:vit_ -c N -o gpurun_out/x python tools/prof_kernels.py ref iq
names: ref (headline), iq (raw-I/Q feeder), lin (textbook labels), r13 (rate 1/3), k9 (tile kernel forward + traceback), hard"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import vdec_pkg
    cc = vdec_pkg.load()
    C = cc.codes
    F = int(os.environ.get("PROF_FRAMES", 148 * 4 * 128 * 8))          # 8 full waves of the register kernels
    for name in sys.argv[1:] or ["ref"]:
        if name in ("ref", "iq", "lin"):
            dec = cc.BatchDecoder(C.K7_R12, cc.SOFT_F32, cc.PARITY_LINEAR if name == "lin" else cc.PARITY_REF)
            if name == "iq":
                _, x = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(4.0), want_iq=True, want_metrics=False)
                dec.decode_iq(x)
            else:
                _, x = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(4.0))
                dec.decode(x)
        elif name == "r13":
            dec = cc.BatchDecoder(C.K7_R13, cc.SOFT_F32, cc.PARITY_REF)
            _, x = dec.synth_awgn(F // 2, 0, 1, cc.awgn_sigma(4.0))
            dec.decode(x)
        elif name == "k9":
            dec = cc.BatchDecoder(C.K9_R12, cc.SOFT_F32, cc.PARITY_LINEAR)
            _, x = dec.synth_awgn(F // 8, 0, 1, cc.awgn_sigma(4.0))
            dec.decode(x)
        elif name == "hard":
            dec = cc.BatchDecoder(C.K7_R12, cc.HARD_U16, cc.PARITY_REF)
            _, x = dec.synth_bsc(F, 0, 1, 20000)
            dec.decode(x)
        else:
            raise SystemExit("unknown kernel name " + name)
        torch.cuda.synchronize()
        print(name, dec.kernel_name, flush=True)
        dec.close()
        del x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
