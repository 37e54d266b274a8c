#!/bin/bash
# One 8-GPU box call: the bench at N = 8 (with and without CPU affinity), config 5 through the C sweep entry.
TAG=${1:-r2n8}; N=${2:-8}; O=gpurun_out; mkdir -p $O results
export VDEC_JIT_CACHE=$PWD/build/jit_cache
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
nvidia-smi --query-gpu=index,name,clocks.max.sm,pci.bus_id --format=csv >> $O/${TAG}_topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $O/${TAG}_topo.txt 2>&1
run_bench() { # name, extra env
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_$1.json 2> $O/${TAG}_bench_$1.err
  echo "bench $1 rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$1.json"))
    print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, {k:(d[k]["value"], d[k].get("h2d_gbs")) for k in ("e2e","e2e_iq","e2e_hard")}, "h2d_peak", d["e2e"]["h2d_peak_gbs"], d["e2e"]["cpu_affinity"], "link", d["link"]["value"])
except Exception as e: print("no json", e)
PY
}
run_bench aff "X=1"
run_bench noaff "BENCH_NO_AFFINITY=1"
# BASELINE config 5: Eb/N0 0..8 dB, ~1e11 bits at the high-SNR points, frames sharded over all GPUs, host code in C
L=convolutional-codes_b200/lib
timeout 900 $L/vdec-sweep --code k7r12 --labels linear --ebn0 0:8:1 --frames 8000000 --min-frame-errors 200 --max-frames 100000000 \
    --out results/awgn_k7r12_linear_soft_b200_r02.m > results/awgn_k7r12_linear_soft_b200_r02.log 2>&1; echo "sweep rc=$?"; tail -4 results/awgn_k7r12_linear_soft_b200_r02.log
timeout 600 $L/vdec-sweep --code k7r12 --labels ref --ebn0 0:8:2 --frames 8000000 --out results/awgn_k7r12_ref_soft_b200_r02.m > results/awgn_k7r12_ref_soft_b200_r02.log 2>&1; tail -2 results/awgn_k7r12_ref_soft_b200_r02.log
timeout 600 $L/vdec-sweep --code cb0 --channel bsc --frames 20000000 --out results/bsc_codebook_b200_r02.m > results/bsc_cb0_b200_r02.log 2>&1; tail -2 results/bsc_cb0_b200_r02.log
# same sweep on 1 GPU, fewer frames: counters must not depend on the GPU count
$L/vdec-sweep --code cb3 --ebn0 2:6:2 --frames 400000 --gpus 1 | grep -E "^bits" > $O/${TAG}_sweep_1gpu.txt
$L/vdec-sweep --code cb3 --ebn0 2:6:2 --frames 400000 --gpus $N | grep -E "^bits" > $O/${TAG}_sweep_ngpu.txt
cmp $O/${TAG}_sweep_1gpu.txt $O/${TAG}_sweep_ngpu.txt && echo "sweep counters identical on 1 and $N GPUs"
cp results/*_r02.* $O/ 2>/dev/null
true
