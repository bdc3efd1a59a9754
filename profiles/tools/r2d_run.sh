#!/bin/bash
# round 2, GPU call D: FMA/SFU kernel with pre-duplicated operands, look-ahead LU probes, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2d_pytest.log | tail -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs-table --factor-sizes=256,1024,2048 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2d_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"])
print("e2e", j.get("e2e"))
print("factor", json.dumps(j.get("factor_ms_by_n")))
PY
tail -3 gpurun_out/r2d_bench.err
FD_LU_DEBUG=3 timeout 120 python profiles/tools/lu_step_probe.py 256 1024 2048 > gpurun_out/r2d_lu_probe.log 2>&1; cat gpurun_out/r2d_lu_probe.log
FD_LU_DEBUG=3 FD_LU_NOLA=1 timeout 120 python profiles/tools/lu_step_probe.py 256 1024 > gpurun_out/r2d_lu_probe_nola.log 2>&1; cat gpurun_out/r2d_lu_probe_nola.log
timeout 300 python profiles/tools/configs_probe.py --only C3g,C1 > gpurun_out/r2d_configs.jsonl 2>&1; cat gpurun_out/r2d_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval64_mma -s 8 -c 1 -o gpurun_out/r2d_eval64_mma -f python profiles/tools/eval64_probe.py > gpurun_out/r2d_ncu_eval64.log 2>&1; echo "ncu eval64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_f32x2 -s 4 -c 1 -o gpurun_out/r2d_eval_f32x2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes= > gpurun_out/r2d_ncu_f32x2.log 2>&1; echo "ncu f32x2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_fused_la -s 4 -c 1 -o gpurun_out/r2d_lu_la -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes= > gpurun_out/r2d_ncu_lu.log 2>&1; echo "ncu lu rc=$?"
