#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): the multi-device tests, the weak-scaling bench line with the NCCL broadcast in
# the timed region, C5 (16M vertices x 4096 control points x 1000 frames, strong scaling) with FP32 forced and under AUTO
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "mgpu or two_contexts or torchrun" > gpurun_out/r2e_pytest_$N.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2e_pytest_$N.log | tail -12
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2e_scale_c2_$N.json 2> gpurun_out/r2e_scale_c2_$N.err; echo "C2 rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --eval-precision 1 > gpurun_out/r2e_scale_c2_fp32_$N.json 2> gpurun_out/r2e_scale_c2_fp32_$N.err; echo "C2 fp32 rc=$?"
timeout 900 $TR bench.py --gpus $N --config C5 --steps 2 --warmup 3 --eval-precision 1 --no-e2e > gpurun_out/r2e_c5_fp32_$N.json 2> gpurun_out/r2e_c5_fp32_$N.err; echo "C5 fp32 rc=$?"
timeout 1500 $TR bench.py --gpus $N --config C5 --steps 1 --warmup 3 --no-e2e > gpurun_out/r2e_c5_auto_$N.json 2> gpurun_out/r2e_c5_auto_$N.err; echo "C5 auto rc=$?"
python - <<PY
import json
for f in ("r2e_scale_c2_$N", "r2e_scale_c2_fp32_$N", "r2e_c5_fp32_$N", "r2e_c5_auto_$N"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: j.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "dtype")}, j["config"]["eval_kernel"], j.get("comm"), j.get("e2e", {}).get("ms_per_step"), j.get("e2e", {}).get("copy_only_ms_per_step"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -3 gpurun_out/r2e_c5_auto_$N.err
