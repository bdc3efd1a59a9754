#!/bin/bash
# round 2, GPU call B: tests, bench line (symmetric DMMA LU, recalibrated AUTO), accuracy calibration
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2b_pytest.log | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2b_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"])
print("e2e", j.get("e2e"))
print("factor", json.dumps(j.get("factor_ms_by_n")))
print("cpu", j.get("cpu_baseline"), j.get("cpu_baseline_1thread"))
for k, v in (j.get("other_configs") or {}).items(): print(k, v)
PY
tail -3 gpurun_out/r2b_bench.err
timeout 600 python tests/tools/accuracy_probe.py 256 1024 2048 4096 > gpurun_out/r2b_accuracy.log 2>&1; echo "accuracy rc=$?"; cat gpurun_out/r2b_accuracy.log
grep -h "AUTO kernel\|path=" gpurun_out/r2b_pytest.log | head -30
timeout 300 python profiles/tools/percook_probe.py > gpurun_out/r2b_percook.jsonl 2> gpurun_out/r2b_percook.err; echo "percook rc=$?"; cat gpurun_out/r2b_percook.jsonl; tail -3 gpurun_out/r2b_percook.err
