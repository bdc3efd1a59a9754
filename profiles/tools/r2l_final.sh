#!/bin/bash
# round 2, last evidence run on one B200 with the shipped defaults (exact-digit kernel: two column blocks per unit, N = 240 MMAs):
# GPU tests, the bench line, ncu --set full of the dominant kernel, the launch list of the fast bench command.
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2l_pytest.log
timeout 120 python bench.py --steps 20 --warmup 3 --cpu-seconds 8 > gpurun_out/r2l_bench_1gpu.json 2> gpurun_out/r2l_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r2l_bench_1gpu.json"))
    print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches")})
    print(j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"], "e2e", j.get("e2e", {}).get("ms_per_step"))
except Exception as e:
    print("bench json:", e)
PY
BENCH_FAST="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes="
timeout 50 ncu --set full --clock-control none --import-source on -k regex:k_eval_tcx -s 4 -c 1 -o gpurun_out/r2l_eval_tcx -f $BENCH_FAST > gpurun_out/r2l_ncu_a.log 2>&1; echo "ncu tcx rc=$?"
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv $BENCH_FAST > gpurun_out/r2l_ncu_launches.log 2>&1; echo "launch list rc=$?"
