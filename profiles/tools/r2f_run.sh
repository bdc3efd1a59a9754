#!/bin/bash
# round 2, multi-GPU call after the exact-digit tensor-core kernel became FD_EVAL_AUTO's wide-batch path: the lines that changed
# (C2 and C5 under AUTO); the multi-device tests again when asked for (second argument "tests").  FP32-forced lines: r2e_run.sh.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "mgpu or two_contexts or torchrun" > gpurun_out/r2f_pytest_$N.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2f_pytest_$N.log | tail -6
fi
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2f_scale_c2_$N.json 2> gpurun_out/r2f_scale_c2_$N.err; echo "C2 rc=$?"
timeout 900 $TR bench.py --gpus $N --config C5 --steps 2 --warmup 2 --no-e2e > gpurun_out/r2f_c5_auto_$N.json 2> gpurun_out/r2f_c5_auto_$N.err; echo "C5 auto rc=$?"
python - <<PY
import json
for f in ("r2f_scale_c2_$N", "r2f_c5_auto_$N"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: j.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "dtype")}, j["config"]["eval_kernel"], j.get("comm", {}).get("ms_per_pass"), j.get("e2e", {}).get("ms_per_step"), j.get("e2e", {}).get("copy_only_ms_per_step"))
    except Exception as ex:
        print(f, "failed", ex)
PY
