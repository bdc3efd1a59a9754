#!/bin/bash
# round 2, GPU call G: warp-specialised FP64 evaluation, 4 vertices per thread in the FMA/SFU kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2g_pytest.log | tail -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs-table --factor-sizes=256 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2g_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"])
PY
tail -3 gpurun_out/r2g_bench.err
timeout 300 python profiles/tools/eval64_probe.py > gpurun_out/r2g_eval64.jsonl 2> gpurun_out/r2g_eval64.err; echo "eval64 rc=$?"; grep -v FP32 gpurun_out/r2g_eval64.jsonl; tail -3 gpurun_out/r2g_eval64.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval64_mma -s 8 -c 1 -o gpurun_out/r2g_eval64_mma -f python profiles/tools/eval64_probe.py > gpurun_out/r2g_ncu_eval64.log 2>&1; echo "ncu eval64 rc=$?"
