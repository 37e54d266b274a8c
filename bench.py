#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: decoded Gbit/s of soft-decision Viterbi,
K=7 rate-1/2 (171,133), 1M frames x 1024 bits per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1).  A step = one pass of the hot path (vdec_decode_device)
over the rank's resident batch of synthetic frames.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "viterbi_decoded_gbit_per_s_k7_r12_soft"
UNIT = "Gbit/s"
EBN0_DB = 4.0
SEED = 12345


def workload_config(frames_per_gpu: int, n_gpus: int) -> dict:
    return {
        "workload": "awgn-simulation: soft-decision Viterbi, K=7 rate-1/2 (171,133)_8, L=1024 (T=1030 steps, "
                    "4 float32 demapper metrics per step), reference trellis labels",
        "frames_per_gpu": frames_per_gpu, "block_len": 1024, "constr_len": 7, "symlen_out": 2,
        "ebn0_db": EBN0_DB, "labels": "ref", "parallelism": f"frames sharded over {n_gpus} GPU(s), no data-path collective",
        "cache": "inputs (16.5 GB/GPU) and survivor scratch (8.2 GB/GPU) are far larger than the 126 MB L2; no flush needed",
    }


# --------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        self.samples, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:           # no NVML: report that, do not invent numbers
            self.err = repr(e)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self) -> dict:
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b]}


def cpu_decode(orc, ocode, metrics, threads: int):
    """CPU restatement of the reference decoder over `threads` host threads (ctypes drops the GIL)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    parts = np.array_split(np.arange(metrics.shape[0]), threads)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda idx: orc.decode_soft(ocode, metrics[idx[0]:idx[-1] + 1]) if len(idx) else None, parts))
    dt = time.perf_counter() - t0
    return dt, outs


def cpu_sample_inputs(orc, ocode, frames: int, seed: int):
    """the same workload generated on the CPU (payload -> encoder -> QPSK -> AWGN -> demapper)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, size=(frames, ocode.payload_bytes), dtype=np.uint8)
    sym = orc.encode(ocode, payload)
    sigma = (2.0 ** -0.5) * 10.0 ** (-EBN0_DB / 20.0)
    iq = orc.constellation(ocode.n)[sym] + (sigma * rng.standard_normal((frames, ocode.T, 2))).astype(np.float32)
    return orc.demap(ocode.n, iq.astype(np.float32).reshape(-1, 2)).reshape(frames, ocode.T, ocode.M)


# --------------------------------------------------------------------------------------------------
def run_reference(args) -> int:
    """--impl reference: the reference's CPU decoder on the host cores, same metric and config.
    The unmodified objects cannot take L=1024 (uint8_t lengths), so this runs the type-widened
    restatement oracle/vit_oracle.c ("port"), pinned bit-for-bit to the unmodified code on K<=8, L<=249."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    orc = oracle.Oracle()
    oc = oracle.Code(2, 7, 1024, [0xF200000000000000, 0xB600000000000000], oracle.PARITY_REF)
    cores = os.cpu_count() or 1
    frames = 512 * cores                     # ~1.5 s of work per core per step
    met = cpu_sample_inputs(orc, oc, frames, SEED)
    for _ in range(args.warmup):
        cpu_decode(orc, oc, met, cores)
    total = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_decode(orc, oc, met, cores)
        total += dt
    gbit = frames * oc.L * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbit, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(1_000_000, args.gpus),
        "cpu_baseline": {"value": gbit, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{frames} frames of the workload per step ({frames * oc.L} decoded bits), all host threads"},
        "e2e": {"value": gbit, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "acs_gstate_updates_per_s": gbit * 64 * 1030 / 1024,
    }
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
def run_gpu(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist
    import vdec_pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cc = vdec_pkg.load()
    code = cc.codes.K7_R12                                   # L = 1024
    F = args.frames
    dec = cc.BatchDecoder(code, cc.SOFT_F32, cc.PARITY_REF, device=local)
    sigma = cc.awgn_sigma(EBN0_DB)
    payload, metrics = dec.synth_awgn(F, rank * F, SEED, sigma)   # resident in HBM before timing
    out = torch.empty((F, dec.out_bytes), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        dec.decode(metrics, out=out)
    barrier()
    l0 = dec.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            dec.decode(metrics, out=out)                  # vdec_decode_device on torch's current stream
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = dec.launches - l0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    bits = world * F * code.L
    gbit = bits / (ms_step * 1e-3) / 1e9

    # error counters of the last step, summed over ranks: the path's one collective (main.c:129-141,233)
    counters = dec.count_errors(out, payload)
    cc.allreduce_counters(counters)
    counters = counters.cpu().tolist()

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernels + D2H in the timed region
    Fe = min(args.e2e_frames, F)
    host_in = torch.empty((Fe, code.T, code.M), dtype=torch.float32, pin_memory=True)
    host_in.copy_(metrics[:Fe])
    host_out = torch.empty((Fe, dec.out_bytes), dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        dec.decode_host(host_in, out=host_out)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dec.decode_host(host_in, out=host_out)            # synchronous: returns after the D2H copy
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_gbit = world * Fe * code.L * e2e_steps / float(te.item()) / 1e9
    assert torch.equal(host_out.to(dev), out[:Fe]), "host-buffer path and device path disagree"

    # ---- same, through the raw-I/Q entry (demapper fused behind the call: half the PCIe bytes per bit)
    _, _, iq_dev = dec.synth_awgn(Fe, rank * F, SEED, sigma, want_iq=True)
    host_iq = torch.empty((Fe, code.T, 2), dtype=torch.float32, pin_memory=True)
    host_iq.copy_(iq_dev)
    del iq_dev
    for _ in range(2):
        dec.decode_iq_host(host_iq, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dec.decode_iq_host(host_iq, out=host_out)
    tq = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
    e2e_iq_gbit = world * Fe * code.L * e2e_steps / float(tq.item()) / 1e9
    assert torch.equal(host_out.to(dev), out[:Fe]), "raw-I/Q path and metric path disagree"

    # ---- the hard-decision decoder end to end (binary-symmetric-channel/viterbi-decoder.c: one symbol byte per step):
    #      1030 B in, 129 B out per frame, so PCIe carries 8 x fewer bytes per decoded bit than the float metrics
    hdec = cc.BatchDecoder(code, cc.HARD_U16, cc.PARITY_REF, device=local)
    Fh = min(4 * Fe, F)
    _, hsym = hdec.synth_bsc(Fh, rank * F, SEED, 20000)
    host_sym = torch.empty((Fh, code.T), dtype=torch.uint8, pin_memory=True)
    host_sym.copy_(hsym)
    host_hout = torch.empty((Fh, hdec.out_bytes), dtype=torch.uint8, pin_memory=True)
    host_hpm = torch.empty((Fh,), dtype=torch.int32, pin_memory=True)
    for _ in range(2):
        hdec.decode_host(host_sym, out=host_hout, path_metric=host_hpm)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hdec.decode_host(host_sym, out=host_hout, path_metric=host_hpm)
    th = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    e2e_hard_gbit = world * Fh * code.L * e2e_steps / float(th.item()) / 1e9
    hout_dev, _ = hdec.decode(hsym)
    assert torch.equal(host_hout.to(dev), hout_dev), "hard-decision host path and device path disagree"
    hard_kernel = hdec.kernel_name
    del hsym, hout_dev
    hdec.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (one launch per step: the register Viterbi kernel)
    hbm_peak, sm_max_mhz, peak_kind = measured_peaks()
    algo_bytes = F * (code.T * code.M * 4 + dec.out_bytes)              # 16.22 B per decoded bit
    kern_ms = ms_step * args.steps / max(launches, 1)                    # launches inside the timed region
    achieved = algo_bytes / (ms_step * 1e-3) / 1e9                       # all launches of a step run the same kernel
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_frame")
        traffic = traffic * F if traffic is not None else None
    clocks = clk.summary()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    state_updates = F * code.T * code.S / (ms_step * 1e-3)              # per GPU
    f_clk = (clocks["sm_mhz"] or sm_max_mhz) * 1e6
    alu = {"unit": "state-updates/s", "achieved": state_updates,
           "issue_peak_at_max_clock": sms * 4 * 32 * sm_max_mhz * 1e6 / 4.0,       # 4 issue slots per update (FADD2, compare, min, decision; 3.1 measured since the sign-based decisions)
           "alu_pipe_peak_at_max_clock": sms * 32 * sm_max_mhz * 1e6 / 1.0,        # 2 half-rate ALU-pipe ops per update
           "frac_of_issue_peak_at_observed_clock": state_updates / (sms * 4 * 32 * f_clk / 4.0)}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on one host core, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        orc = oracle.Oracle()
        oc = oracle.Code(code.n, code.K, code.L, code.polynomials, oracle.PARITY_REF)
        ns = args.cpu_frames
        sample = metrics[:ns].cpu().numpy()
        dt, outs = cpu_decode(orc, oc, sample, 1)
        assert np.array_equal(outs[0], out[:ns].cpu().numpy()), "GPU output differs from the CPU oracle on the sample"
        cpu = {"value": ns * code.L / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {ns} frames of the GPU batch ({dt:.1f} s on one core); GPU output checked bit-exact on them"}
        # and the UNMODIFIED reference objects (oracle/_ref) on the longest frame their uint8_t interface takes
        if oracle.reference_available():
            ref = oracle.Reference()
            c249 = code.with_block_len(249)
            o249 = oracle.Code(c249.n, c249.K, c249.L, c249.polynomials, oracle.PARITY_REF)
            d249 = cc.BatchDecoder(c249, cc.SOFT_F32, cc.PARITY_REF, device=local)
            nf = 2 * ns
            _, m249 = d249.synth_awgn(nf, 0, SEED + 1, sigma)
            g249 = d249.decode(m249).cpu().numpy()
            h249 = m249.cpu().numpy()
            t0 = time.perf_counter()
            r249 = ref.awgn_decode(o249, h249)
            dt2 = time.perf_counter() - t0
            assert np.array_equal(r249, g249), "GPU output differs from the compiled reference decoder"
            cpu["reference_objects"] = {"value": nf * c249.L / dt2 / 1e9, "unit": UNIT, "cores": 1,
                                        "sample": f"{nf} frames of K=7 (171,133) L=249 through AWGN-channel/viterbi-decoder.c compiled "
                                                  f"unmodified ({dt2:.1f} s); GPU output bit-exact on all of them"}
            d249.close()

    line = {
        "metric": METRIC, "value": gbit, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(F, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "kernel": dec.kernel_name,
                     "kernel_ms": kern_ms, "note": "ACS is ALU-bound (min-plus recurrence); see alu_roofline"},
        "alu_roofline": alu,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_gbit, "unit": UNIT, "h2d_bytes_per_step": Fe * dec.in_bytes, "d2h_bytes_per_step": Fe * dec.out_bytes,
                "frames_per_step": Fe, "api": "vdec_decode_host (pinned host buffers, 2-stream pipeline)"},
        "e2e_iq": {"value": e2e_iq_gbit, "unit": UNIT, "h2d_bytes_per_step": Fe * code.T * 8, "d2h_bytes_per_step": Fe * dec.out_bytes,
                   "frames_per_step": Fe, "api": "vdec_decode_iq_host (raw I/Q in, demapper on the GPU; SURVEY 8f-2)"},
        "e2e_hard": {"value": e2e_hard_gbit, "unit": UNIT, "h2d_bytes_per_step": Fh * code.T, "d2h_bytes_per_step": Fh * (dec.out_bytes + 4),
                     "frames_per_step": Fh, "kernel": hard_kernel,
                     "api": "vdec_decode_host, VDEC_HARD_U16 (the BSC decoder's input: one symbol byte per step), same code and frame length"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "acs_gstate_updates_per_s": world * state_updates / 1e9,
        "bit_errors": {"bits": counters[0], "bit_errors": counters[1], "frame_errors": counters[2]},
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line: dict) -> None:
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) went to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main() -> int:
    os.dup2(2, 1)                      # libraries that print to stdout (NCCL version banner) must not break the contract
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1_000_000, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=65536)
    ap.add_argument("--cpu-frames", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
