#!/bin/bash
# One GPU box call: tests, bench (both arms), ncu launch list + full captures, pipe microbenchmark, sanitizers, A/B builds.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r2a "tests bench ncu pipes sanitize ab"'
TAG=${1:-r2x}; WHAT=${2:-"tests bench"}; O=gpurun_out; mkdir -p $O
export VDEC_JIT_CACHE=$PWD/build/jit_cache
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then timeout 1200 python -m pytest tests -q -m gpu -x --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -5 $O/${TAG}_pytest.log; fi
if has alltests; then timeout 1500 python -m pytest tests -q -m gpu --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -15 $O/${TAG}_pytest.log; fi
if has smoke; then timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log; fi
if has bench; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
  timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; head -c 1500 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_regs_kernel -c 2 -f -o $O/${TAG}_full_ref_iq \
      python tools/prof_kernels.py ref iq > $O/${TAG}_ncu_full1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_regs_kernel -c 2 -f -o $O/${TAG}_full_lin_r13 \
      python tools/prof_kernels.py lin r13 > $O/${TAG}_ncu_full2.log 2>&1
  ls -la $O/*.ncu-rep
fi
if has pipes; then ./build/pipes > $O/${TAG}_pipes.jsonl 2>&1; tail -3 $O/${TAG}_pipes.jsonl; fi
if has sanitize; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/${TAG}_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_run.py > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/${TAG}_racecheck.log
fi
if has ab; then
  for v in ${AB_VARIANTS:-base}; do
    d=convolutional-codes_b200/lib; [[ $v != base ]] && d=build/$v
    for c in ${AB_CASES:-"C2 soft K7 r1/2 L1024 linear"}; do :; done
    VDEC_LIB_DIR=$PWD/$d timeout 300 python tools/bench_configs.py --only "${AB_ONLY:-C2 soft K7}" > $O/${TAG}_ab_$v.jsonl 2> $O/${TAG}_ab_$v.err
    echo "== $v"; cat $O/${TAG}_ab_$v.jsonl
  done
fi
true
