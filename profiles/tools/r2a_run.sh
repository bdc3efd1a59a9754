#!/bin/bash
# round 2, GPU call A: tests, bench line, accuracy calibration, FP64 evaluation timings, FP64 micro-benchmark
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2a_bench.json
timeout 600 python tests/tools/accuracy_probe.py 256 1024 2048 4096 > gpurun_out/r2a_accuracy.log 2>&1; echo "accuracy rc=$?"; cat gpurun_out/r2a_accuracy.log
timeout 300 python profiles/tools/eval64_probe.py > gpurun_out/r2a_eval64.jsonl 2> gpurun_out/r2a_eval64.err; echo "eval64 rc=$?"; cat gpurun_out/r2a_eval64.jsonl; tail -3 gpurun_out/r2a_eval64.err
(cd profiles/tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu && ./fp64_tput) > gpurun_out/r2a_fp64_tput.log 2>&1; cat gpurun_out/r2a_fp64_tput.log
