#!/bin/bash
# round 2, variants of the exact-digit kernel on one B200: default (two rows per producer thread + N = 240 MMAs), each
# switched off in turn, and the configuration of commit a6c0a9c (both off).  Accuracy gate (tcx_check) + fast bench for each,
# the GPU test suite under the default.
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 60 python profiles/tools/tcx_check.py > gpurun_out/r2k_tcx_check_$name.log 2>&1
  echo "== $name: $(grep -c 'max' gpurun_out/r2k_tcx_check_$name.log) cases, worst $(grep 'max' gpurun_out/r2k_tcx_check_$name.log | sed 's/.*diag = \([0-9.e+-]*\).*/\1/' | sort -g | tail -1), nonfinite/err: $(grep -ci 'false\|error\|traceback' gpurun_out/r2k_tcx_check_$name.log)"
  env "$@" timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes= > gpurun_out/r2k_bench_fast_$name.json 2> gpurun_out/r2k_bench_fast_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r2k_bench_fast_$name.json')); print('   step', round(j['ms_per_step'],4), 'eval', round(j['phase_ms_last_step']['eval'],4), 'launch', round(j['roofline']['launch_ms'],4))" 2>&1 | tail -1
}
run default FD_DUMMY=0
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_default.log 2>&1; tail -2 gpurun_out/r2k_pytest_default.log
run rows2_narrow FD_TCX_NARROW=1
run rows1_wide FD_TCX_ROWS=1
run rows1_narrow FD_TCX_ROWS=1 FD_TCX_NARROW=1
