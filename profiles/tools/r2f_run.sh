#!/bin/bash
# round 2, GPU call F: column-packed FMA/SFU kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2f_pytest.log | tail -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs-table --factor-sizes=256 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2f_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"])
print("e2e", j.get("e2e"))
PY
tail -3 gpurun_out/r2f_bench.err
timeout 300 python profiles/tools/configs_probe.py --only C3g,C1 > gpurun_out/r2f_configs.jsonl 2>&1; cat gpurun_out/r2f_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_f32c -s 4 -c 1 -o gpurun_out/r2f_eval_f32c -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes= > gpurun_out/r2f_ncu_f32c.log 2>&1; echo "ncu f32c rc=$?"
