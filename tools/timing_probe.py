import sys, os, json, time
sys.path.insert(0, os.getcwd())
import torch, vdec_pkg
cc = vdec_pkg.load()
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
code = cc.codes.K7_R12
F = 1_000_000
dec = cc.BatchDecoder(code, cc.SOFT_F32, cc.PARITY_REF)
payload, metrics = dec.synth_awgn(F, 0, 12345, cc.awgn_sigma(4.0))
out = torch.empty((F, dec.out_bytes), dtype=torch.uint8, device="cuda")
counters = torch.zeros(3, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
def run(label, n, with_count):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    clk = []
    evs[0].record()
    for i in range(n):
        dec.decode(metrics, out=out)
        if with_count:
            counters.zero_(); dec.count_errors(out, payload, counters)
        evs[i + 1].record()
    for i in range(n):
        evs[i + 1].synchronize()
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    ts = [round(evs[i].elapsed_time(evs[i + 1]), 3) for i in range(n)]
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000
    print(label, ts, "clk", clk[-3:], "power", pw, flush=True)
run("warm", 3, False)
run("decode x13", 13, False)
run("decode+count x13", 13, True)
run("decode x13", 13, False)
time.sleep(2)
run("after 2 s idle: decode x6", 6, False)
# different buffer offsets
m2 = torch.empty((F * code.T * code.M + 4096,), dtype=torch.float32, device="cuda")
for off in (0, 32, 1024):
    v = m2[off:off + F * code.T * code.M].view(F, code.T, code.M); v.copy_(metrics)
    mm = metrics; metrics = v
    run(f"input offset {off*4} B", 5, False)
    metrics = mm
for Fx in (999_424, 1_000_000, 1_048_576, 909_312):    # 999424 = 7808 CTAs; 909312 = 12 waves exactly
    mv = metrics[:Fx]; ov = out[:Fx]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    if Fx > F: continue
    evs[0].record()
    for i in range(5):
        dec.decode(mv, out=ov); evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(5)]
    print("F", Fx, "CTAs", (Fx + 127) // 128, "waves", round((Fx + 127) // 128 / 592, 2), "gbit/s", round(Fx * 1024 / (sum(ts[1:]) / 4) / 1e6, 2), flush=True)
