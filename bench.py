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
# the per-step counter all-reduce: in line on the compute stream (default) or overlapped on NCCL's stream (A/B switch)
SYNC_ALLREDUCE = os.environ.get("BENCH_ASYNC_ALLREDUCE", "0") == "0"


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
def set_cpu_affinity(local: int) -> str:
    """Bind this rank to the CPUs next to its GPU before pinned host memory is first touched (NUMA-local staging)."""
    if os.environ.get("BENCH_NO_AFFINITY"):
        return "disabled (BENCH_NO_AFFINITY)"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"nvml ideal cpu set ({len(os.sched_getaffinity(0))} of {before} cpus)"
    except Exception as e:
        return "unavailable: " + type(e).__name__


def algorithmic_bytes_per_frame(code, metric_kind: str, kernel: str) -> int:
    """boundary form = input + output per frame; kernels whose survivors live in HBM by design (K >= 8,
    BASELINE config 4) add the survivor stream: S/8 B per step written, one 32 B sector per step read back."""
    step = {"soft": code.M * 4, "iq": 8, "hard": 1, "quant": code.M}[metric_kind]
    b = code.T * step + code.out_bytes + (4 if metric_kind == "hard" else 0)
    if code.K >= 8:
        b += code.T * (code.S // 8) + code.T * 32
    return b


def time_decode(torch, fn, reps: int, warm: int = 2) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_configs(cc, torch, local: int, alu_peak_ops: float, hbm_peak_gbs: float, quick: int):
    """Every other BASELINE.json config shape, device-resident, timed in the same run: decoded Gbit/s, ACS
    Gstate-updates/s and both roofline fractions per entry.  alu_frac counts 4 lane-operations per state update against
    SMs x 128 lanes x max clock; the packed 16-bit kernels process two lanes per 32-bit lane (ceiling x 2)."""
    C = cc.codes
    q = quick
    cases = [  # id, description, code, metric kind, parity, frames, flags
        ("C1", "binary-symmetric-simulation: hard K=3 (7,5) codebook 0, L=40, BSC p=0.0125", cc.get_code(0), "hard", cc.PARITY_REF, (1 << 22) // q, 0),
        ("C2-textbook", "K=7 (171,133) soft, true-parity (textbook) labels, L=1024", C.K7_R12, "soft", cc.PARITY_LINEAR, 500000 // q, 0),
        ("C2-iq", "K=7 (171,133) soft from raw I/Q (demapper in the kernel's feeder), reference labels", C.K7_R12, "iq", cc.PARITY_REF, 1000000 // q, 0),
        ("C2-hard", "K=7 (171,133) hard decisions (BSC p=0.02), packed 16-bit", C.K7_R12, "hard", cc.PARITY_REF, 1000000 // q, 0),
        ("C2-quant", "K=7 (171,133) quantised soft (u8 metrics <= 7), packed 16-bit", C.K7_R12, "quant", cc.PARITY_REF, 1000000 // q, 0),
        ("C3", "K=7 rate-1/3 (133,171,165) soft, 8-point constellation, L=1024, reference labels", C.K7_R13, "soft", cc.PARITY_REF, 500000 // q, 0),
        ("C3-textbook", "K=7 rate-1/3 soft, true-parity labels", C.K7_R13, "soft", cc.PARITY_LINEAR, 500000 // q, 0),
        ("C3-iq", "K=7 rate-1/3 from raw I/Q: the simulation's soft demapper (common/demapper.c) inside the decoder kernel's feeder", C.K7_R13, "iq", cc.PARITY_REF, 500000 // q, 0),
        ("C4-K9", "K=9 (561,753) soft, L=1024, survivors in HBM, true-parity labels", C.K9_R12, "soft", cc.PARITY_LINEAR, 131072 // q, 0),
        ("C4-K9-ref", "K=9 (561,753) soft, reference labels", C.K9_R12, "soft", cc.PARITY_REF, 131072 // q, 0),
        ("C4-K15", "K=15 (46321,51271) soft, 16384 states, L=1024, survivors in HBM, true-parity labels", C.K15_R12, "soft", cc.PARITY_LINEAR, 2368 // min(q, 2), 0),
        ("C4-K15-ref", "K=15 soft, reference labels", C.K15_R12, "soft", cc.PARITY_REF, 2368 // min(q, 2), 0),
    ]
    out = []
    dev = torch.device("cuda", local)
    for cid, desc, code, kind, parity, F, flags in cases:
        metric = {"soft": cc.SOFT_F32, "iq": cc.SOFT_F32, "hard": cc.HARD_U16, "quant": cc.SOFT_U8}[kind]
        dec = cc.BatchDecoder(code, metric, parity, device=local, flags=flags, metric_max=7 if kind == "quant" else 0)
        if kind == "soft":
            payload, x = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(EBN0_DB))
        elif kind == "iq":
            payload, x = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(EBN0_DB), want_iq=True, want_metrics=False)
        elif kind == "quant":
            payload, xf = dec.synth_awgn(F, 0, 1, cc.awgn_sigma(EBN0_DB))
            x = torch.clamp(torch.round(xf * 2.0), max=7).to(torch.uint8)
            del xf
        else:
            payload, x = dec.synth_bsc(F, 0, 1, 12500 if cid == "C1" else 20000)
        res = torch.empty((F, dec.out_bytes), dtype=torch.uint8, device=dev)
        pm = torch.empty((F,), dtype=torch.int32, device=dev)
        fn = (lambda: dec.decode_iq(x, out=res)) if kind == "iq" else (lambda: dec.decode(x, out=res, path_metric=pm))
        l0 = dec.launches
        ms = time_decode(torch, fn, reps=5)
        launches_per_step = (dec.launches - l0) // 7
        cnt = dec.count_errors(res, payload).cpu().tolist()
        kernel = dec.iq_kernel_name if kind == "iq" else dec.kernel_name
        packed = kernel.startswith(("regs2_", "tile2_"))
        upd = F * code.T * code.S / (ms * 1e-3)
        algo = F * algorithmic_bytes_per_frame(code, kind, kernel)
        alu_frac = upd * 4.0 / (alu_peak_ops * (2.0 if packed else 1.0))
        hbm_frac = algo / (ms * 1e-3) / 1e9 / hbm_peak_gbs
        out.append({"id": cid, "workload": desc, "kernel": kernel, "frames": F, "ms": round(ms, 4),
                    "gbit_s": round(F * code.L / ms / 1e6, 3), "gstate_upd_s": round(upd / 1e9, 1),
                    "alu_frac": round(alu_frac, 4), "hbm_frac": round(hbm_frac, 4), "bound": "alu" if alu_frac >= hbm_frac else "hbm",
                    "algorithmic_bytes_per_frame": algo // F, "launches_per_step": launches_per_step,
                    "ber": cnt[1] / max(cnt[0], 1)})
        dec.close()
        del x, res, pm, payload
        torch.cuda.empty_cache()
    # sliding-window decoding of one unterminated stream (north_star piece 4)
    code = C.K7_R12
    dec = cc.BatchDecoder(code, cc.SOFT_F32, cc.PARITY_LINEAR, device=local)
    nb = ((1 << 28) // q) // code.T
    _, x = dec.synth_awgn(nb, 0, 1, cc.awgn_sigma(EBN0_DB))
    x = x.reshape(nb * code.T, code.M)
    ms = time_decode(torch, lambda: dec.decode_stream(x, 1024, 128, start_known=True), reps=3, warm=1)
    steps = x.shape[0]
    upd = (steps / 1024) * 1280 * code.S / (ms * 1e-3)          # every window of 1024 payload steps runs 1280 trellis steps
    out.append({"id": "stream", "workload": "one unterminated K=7 soft stream of 2^28 steps, windows of 1024 + 2 x 128 steps, stitched",
                "kernel": dec.kernel_name, "steps": steps, "ms": round(ms, 4), "gbit_s": round(steps / ms / 1e6, 3),
                "gstate_upd_s": round(upd / 1e9, 1), "alu_frac": round(upd * 4.0 / alu_peak_ops, 4),
                "hbm_frac": round(steps * 1.25 * 16 / (ms * 1e-3) / 1e9 / hbm_peak_gbs, 4), "bound": "alu"})
    dec.close()
    del x
    torch.cuda.empty_cache()
    return out


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
    affinity = set_cpu_affinity(local)
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

    # One step of the sharded simulation (AWGN-channel/main.c:108-141 for a whole batch): decode the rank's frames,
    # count bit / frame errors against the payload, and sum the three counters over the ranks — the path's ONE
    # collective (NCCL all-reduce of 24 bytes), inside the timed region.
    kev = []                                              # (start, end) events around the decoder launch of every timed step
    # one counter vector per step: the all-reduce of step k runs on NCCL's stream while the rank already decodes step
    # k + 1 (its result is only needed at the end), so the ranks are not forced into lock-step by a 24-byte collective
    nsteps = max(args.warmup, 0) + args.steps
    step_counters = torch.zeros((nsteps, 3), dtype=torch.int64, device=dev)
    works = []

    def step(i, timed=False):
        cnt = step_counters[i]
        if timed:
            kev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            kev[-1][0].record()
        dec.decode(metrics, out=out)                      # vdec_decode_device on torch's current stream
        if timed:
            kev[-1][1].record()
        dec.count_errors(out, payload, cnt)
        if world > 1:
            if SYNC_ALLREDUCE:
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            else:
                works.append(dist.all_reduce(cnt, op=dist.ReduceOp.SUM, async_op=True))

    for i in range(max(args.warmup, 0)):
        step(i)
    for w in works:
        w.wait()
    works.clear()
    l0 = dec.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)                             # NVML start-up (tens of ms, different on every rank) stays outside
    with clk:
        barrier()                                         # the ranks enter the timed region together: a late rank would be
        e0.record()                                       # waited for in the first all-reduce and billed to all the others
        for i in range(args.steps):
            step(max(args.warmup, 0) + i, timed=True)
        for w in works:                                   # every step's collective has completed before the clock stops
            w.wait()
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = dec.launches - l0
    # the decoder kernel alone: CUDA events on the launching stream around its launch inside the very same timed steps
    # (the SM clock of this part steps down from 1965 MHz within ~60 ms of sustained load, so a separate loop would not
    # measure the same kernel time)
    dec_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = torch.tensor([dec_ms, ms_total / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [round(float(x[0]), 3) for x in allr], "ms_per_step": [round(float(x[1]), 3) for x in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    bits = world * F * code.L
    gbit = bits / (ms_step * 1e-3) / 1e9
    totals = step_counters[-1].cpu().tolist()             # all ranks, last step
    assert all(torch.equal(step_counters[i], step_counters[-1]) for i in range(nsteps)), "steps disagree"

    if args.value_only:
        if rank == 0:
            _emit({"n_gpus": world, "value": gbit, "ms_per_step": ms_step, "per_rank": per_rank,
                   "allreduce": "sync" if SYNC_ALLREDUCE else "async"})
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- GPU-resident link (config 5's inner loop): source -> channel (raw I/Q) -> demapper+decoder -> error count
    Fl = min(F, 1 << 19)
    lcount = torch.zeros(3, dtype=torch.int64, device=dev)
    lout = torch.empty((Fl, dec.out_bytes), dtype=torch.uint8, device=dev)

    def link_step():
        pl, iq = dec.synth_awgn(Fl, rank * F, SEED + 1, sigma, want_iq=True, want_metrics=False)
        dec.decode_iq(iq, out=lout)
        dec.count_errors(lout, pl, lcount)

    link_ms = time_decode(torch, link_step, reps=3, warm=1)
    tl = torch.tensor([link_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    link_gbit = world * Fl * code.L / (float(tl.item()) * 1e-3) / 1e9
    iq_kernel = dec.iq_kernel_name

    # ---- bare pinned host -> device bandwidth, all ranks at once: the ceiling of every host-buffer number below
    def h2d_peak():
        hb = torch.empty((256 << 20,), dtype=torch.uint8, pin_memory=True)
        hb.fill_(1)                                              # first touch on this rank's CPUs
        db = torch.empty_like(hb, device=dev)
        best = None
        for _ in range(3):                                       # best of three series (the first one after a kernel phase can run slow)
            db.copy_(hb, non_blocking=True)
            barrier()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(8):
                db.copy_(hb, non_blocking=True)                 # cudaMemcpyAsync, 256 MiB each
            h1.record()
            torch.cuda.synchronize()
            th2d = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(th2d, op=dist.ReduceOp.MAX)
            gbs = world * 8 * hb.numel() / (float(th2d.item()) * 1e-3) / 1e9
            best = gbs if best is None else max(best, gbs)
        return best

    h2d_peak_gbs = h2d_peak()

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
    e2e_h2d_gbs = world * Fe * dec.in_bytes * e2e_steps / float(te.item()) / 1e9
    assert torch.equal(host_out.to(dev), out[:Fe]), "host-buffer path and device path disagree"

    # ---- same, through the raw-I/Q entry (demapper inside the decoder kernel: half the PCIe bytes per bit)
    _, iq_dev = dec.synth_awgn(Fe, rank * F, SEED, sigma, want_iq=True, want_metrics=False)
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
    e2e_iq_h2d_gbs = world * Fe * code.T * 8 * e2e_steps / float(tq.item()) / 1e9
    assert torch.equal(host_out.to(dev), out[:Fe]), "raw-I/Q path and metric path disagree"

    # ---- the hard-decision decoder end to end (binary-symmetric-channel/viterbi-decoder.c: one symbol byte per step):
    #      1030 B in, 129 B out per frame, so PCIe carries 8 x fewer bytes per decoded bit than the float metrics
    hdec = cc.BatchDecoder(code, cc.HARD_U16, cc.PARITY_REF, device=local)
    Fh = min(8 * Fe, F)
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
    e2e_hard_h2d_gbs = world * Fh * code.T * e2e_steps / float(th.item()) / 1e9
    hout_dev, _ = hdec.decode(hsym)
    assert torch.equal(host_hout.to(dev), hout_dev), "hard-decision host path and device path disagree"
    hard_kernel = hdec.kernel_name
    del hsym, hout_dev
    hdec.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (one launch per step: the register Viterbi kernel), SURVEY 8d:
    #      achieved = max(ALU term, HBM term), both reported; ACS is a min-plus recurrence, so the ALU term binds
    hbm_peak, sm_max_mhz, peak_kind = measured_peaks()
    clocks = clk.summary()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    alu_peak_ops = sms * 128 * sm_max_mhz * 1e6                          # lane-operations/s at the max SM clock
    f_obs = (clocks["sm_mhz"] or sm_max_mhz) * 1e6
    state_updates = F * code.T * code.S / (dec_ms * 1e-3)               # per GPU, decoder kernel alone
    algo_bytes = F * (code.T * code.M * 4 + dec.out_bytes)              # 16.22 B per decoded bit
    alu_frac = state_updates * 4.0 / alu_peak_ops
    hbm_gbs = algo_bytes / (dec_ms * 1e-3) / 1e9
    hbm_frac = hbm_gbs / hbm_peak
    traffic, traffic_source = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("dram_bytes_per_frame") is not None:
            traffic = tj["dram_bytes_per_frame"] * F
            traffic_source = ("STORED ncu figure, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per frame from "
                              + str(tj.get("source", "profiles/ncu_traffic.json")) + ", times the frames of one launch")
    bound = "alu" if alu_frac >= hbm_frac else "hbm"
    roofline = {
        "bound": bound,
        "achieved": state_updates * 4.0 / 1e12 if bound == "alu" else hbm_gbs,
        "peak": alu_peak_ops / 1e12 if bound == "alu" else hbm_peak,
        "unit": "Tlane-op/s (4 lane-operations per ACS state update: add, add, compare, select)" if bound == "alu" else "GB/s",
        "frac": max(alu_frac, hbm_frac),
        "alu_frac": alu_frac, "alu_frac_at_observed_clock": state_updates * 4.0 / (sms * 128 * f_obs),
        "alu_achieved_state_updates_per_s": state_updates, "alu_peak_state_updates_per_s": alu_peak_ops / 4.0,
        "alu_peak_source": f"{sms} SMs x 128 lanes x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz, {peak_kind})",
        "hbm_frac": hbm_frac, "hbm_achieved_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak,
        "hbm_peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
        "algorithmic_bytes_per_launch": algo_bytes, "traffic": traffic, "traffic_source": traffic_source,
        "kernel": dec.kernel_name, "kernel_ms": dec_ms,
        "kernel_share_of_step": dec_ms / ms_step,
    }

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

    # ---- the other BASELINE configs, same run (N=1 only: they are per-GPU kernels)
    del metrics, payload
    torch.cuda.empty_cache()
    configs = None
    if world == 1 and not args.no_configs:
        configs = run_configs(cc, torch, local, alu_peak_ops, hbm_peak, args.quick)

    line = {
        "metric": METRIC, "value": gbit, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(F, world),
        "step": "decode the rank's frames + count bit/frame errors + NCCL all-reduce of the 3 counters, all inside the timed region",
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_gbit, "unit": UNIT, "h2d_bytes_per_step": Fe * dec.in_bytes, "d2h_bytes_per_step": Fe * dec.out_bytes,
                "frames_per_step": Fe, "api": "vdec_decode_host (pinned host buffers, 2-stream pipeline)",
                "h2d_gbs": e2e_h2d_gbs, "h2d_peak_gbs": h2d_peak_gbs, "frac_of_h2d_peak": e2e_h2d_gbs / h2d_peak_gbs,
                "h2d_peak_how": f"{world} rank(s) at once, 8 x cudaMemcpyAsync of 256 MiB from pinned host memory, nothing else running, best of 3 series",
                "cpu_affinity": affinity},
        "e2e_iq": {"value": e2e_iq_gbit, "unit": UNIT, "h2d_bytes_per_step": Fe * code.T * 8, "d2h_bytes_per_step": Fe * dec.out_bytes,
                   "frames_per_step": Fe, "h2d_gbs": e2e_iq_h2d_gbs, "frac_of_h2d_peak": e2e_iq_h2d_gbs / h2d_peak_gbs, "kernel": iq_kernel,
                   "api": "vdec_decode_iq_host (raw I/Q in, demapper inside the decoder kernel; SURVEY 8f-2)"},
        "e2e_hard": {"value": e2e_hard_gbit, "unit": UNIT, "h2d_bytes_per_step": Fh * code.T, "d2h_bytes_per_step": Fh * (dec.out_bytes + 4),
                     "frames_per_step": Fh, "kernel": hard_kernel, "h2d_gbs": e2e_hard_h2d_gbs, "frac_of_h2d_peak": e2e_hard_h2d_gbs / h2d_peak_gbs,
                     "api": "vdec_decode_host, VDEC_HARD_U16 (the BSC decoder's input: one symbol byte per step), same code and frame length"},
        "link": {"value": link_gbit, "unit": UNIT, "frames_per_step": Fl, "ms_per_step": float(tl.item()), "kernel": iq_kernel,
                 "what": "GPU-resident link step of config 5 (SURVEY 8f-1): synthetic source -> encoder -> mapper -> AWGN (raw I/Q to HBM) -> "
                         "demapper + decoder in one kernel -> error count; nothing touches the host"},
        "per_rank": per_rank,
        "allreduce": "in line on the compute stream" if SYNC_ALLREDUCE else "overlapped on NCCL's stream (async_op), awaited before the end timestamp",
        "gpu_launches": int(launches),
        "clocks": clocks,
        "acs_gstate_updates_per_s": world * F * code.T * code.S / (ms_step * 1e-3) / 1e9,
        "bit_errors": {"bits": totals[0], "bit_errors": totals[1], "frame_errors": totals[2], "ranks_summed": world},
        "configs": configs,
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
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (N=1 only)")
    ap.add_argument("--value-only", action="store_true", help="diagnostic: print per-rank step and kernel times after the timed steps and exit")
    ap.add_argument("--quick", type=int, default=1, help="divide the frame counts of the per-config table by this")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
