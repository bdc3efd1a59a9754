#!/bin/bash
# round 2, GPU call C: look-ahead LU, 3-way FD_EVAL_AUTO, inverse apply v2, DMMA panel GEMM
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|AUTO kernel|path=|radius = spacing" gpurun_out/r2c_pytest.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2c_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"])
print("e2e", j.get("e2e"))
print("factor", json.dumps(j.get("factor_ms_by_n")))
for k, v in (j.get("other_configs") or {}).items(): print(k, v)
PY
tail -3 gpurun_out/r2c_bench.err
FD_LU_NOLA=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e > gpurun_out/r2c_bench_nola.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_nola.json')); print('NOLA', j['phase_ms_last_step'], json.dumps(j['factor_ms_by_n']))"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs-table --factor-sizes= --eval-precision 1 > gpurun_out/r2c_bench_fp32.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_fp32.json')); print('FP32 forced', j['value'], j['ms_per_step'], j['phase_ms_last_step'], j['config']['eval_kernel'])"
timeout 300 python profiles/tools/percook_probe.py > gpurun_out/r2c_percook.jsonl 2> gpurun_out/r2c_percook.err; echo "percook rc=$?"; cat gpurun_out/r2c_percook.jsonl; tail -3 gpurun_out/r2c_percook.err
timeout 300 python profiles/tools/eval64_probe.py > gpurun_out/r2c_eval64.jsonl 2> gpurun_out/r2c_eval64.err; echo "eval64 rc=$?"; cat gpurun_out/r2c_eval64.jsonl
timeout 900 python tests/tools/accuracy_probe.py 256 1024 2048 4096 > gpurun_out/r2c_accuracy.log 2>&1; echo "accuracy rc=$?"; cat gpurun_out/r2c_accuracy.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval64_mma -s 2 -c 1 -o gpurun_out/r2c_eval64_mma -f python profiles/tools/eval64_probe.py > gpurun_out/r2c_ncu_eval64.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c_ncu_eval64.log
