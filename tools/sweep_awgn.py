#!/usr/bin/env python
"""Operating-point sweep on the GPU (BASELINE config 5), one process per GPU: frames sharded over the ranks of one
node, error counters all-reduced per point, results written as JSON lines and as Octave vectors in the style of the
reference's results/awgn_channel.m / binary_symmetric_channel.m.  (The single-process C twin is
convolutional-codes_b200/lib/vdec-sweep, examples/vdec_sweep_main.c.)

    python tools/sweep_awgn.py --code k7r12 --labels linear --ebn0 0:8:1 --frames 200000 --min-frame-errors 200
    python tools/sweep_awgn.py --code cb1 --channel awgn-hard --ebn0 0:14:2        (hard demapper: the *_h curves)
    python tools/sweep_awgn.py --code cb0 --channel bsc                             (the BSC driver's 17 error rates)
    torchrun --standalone --nproc-per-node 8 tools/sweep_awgn.py ...               (one process per GPU)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import vdec_pkg
    cc = vdec_pkg.load()
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="k7r12", choices=["cb0", "cb1", "cb2", "cb3", "cb5", "k7r12", "k7r13", "k9", "k15"])
    ap.add_argument("--labels", default="linear", choices=["ref", "linear"])
    ap.add_argument("--channel", default="awgn", choices=["awgn", "awgn-hard", "bsc"])
    ap.add_argument("--block-len", type=int, default=0, help="0 = the code's default (40 for the codebook, 1024 otherwise)")
    ap.add_argument("--ebn0", default="0:8:1", help="start:stop:step in dB (inclusive)")
    ap.add_argument("--frames", type=int, default=100000, help="frames per batch and point (all ranks together)")
    ap.add_argument("--min-frame-errors", type=int, default=100)
    ap.add_argument("--max-frames", type=int, default=100_000_000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    table = {"cb0": cc.get_code(0), "cb1": cc.get_code(1), "cb2": cc.get_code(2), "cb3": cc.get_code(3), "cb5": cc.get_code(5),
             "k7r12": cc.codes.K7_R12, "k7r13": cc.codes.K7_R13, "k9": cc.codes.K9_R12, "k15": cc.codes.K15_R12}
    code = table[args.code]
    if args.block_len:
        code = code.with_block_len(args.block_len)
    a, b, st = (float(x) for x in args.ebn0.split(":"))
    grid = [a + i * st for i in range(int(round((b - a) / st)) + 1)]
    parity = cc.PARITY_REF if args.labels == "ref" else cc.PARITY_LINEAR
    if args.channel == "bsc":
        dec = cc.BatchDecoder(code, cc.HARD_U16, parity, device=local)
        pts = cc.sweep_bsc(dec, list(cc.BSC_RATES_PPM), args.frames, args.seed, rank, world, min_frame_errors=args.min_frame_errors,
                           max_frames=args.max_frames)
    else:
        dec = cc.BatchDecoder(code, cc.SOFT_F32, parity, device=local)
        pts = cc.sweep_awgn(dec, grid, args.frames, args.seed, rank, world, min_frame_errors=args.min_frame_errors,
                            max_frames=args.max_frames, hard_demap=args.channel == "awgn-hard")
    if rank == 0:
        for p in pts:
            print(json.dumps(dict(p, code=code.name, labels=args.labels, block_len=code.L, kernel=dec.kernel_name, n_gpus=world)))
        if args.out:
            cc.write_octave(args.out, f"{args.code}_{args.labels}_{ {'awgn': 'soft', 'awgn-hard': 'hard', 'bsc': 'bsc'}[args.channel] }", pts)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
