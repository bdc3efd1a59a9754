#!/bin/bash
# round 2, evidence run on one B200: tests, bench line, CPU arm, launch list, ncu --set full of the dominant kernels, probes.
# Outputs land in gpurun_out/r2_*; the summaries kept under profiles/ are made from them (profiles/README.md).
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"
BENCH_FAST="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs-table --no-e2e --factor-sizes="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $BENCH_FAST > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_tcx -s 4 -c 1 -o gpurun_out/r2_eval_tcx -f $BENCH_FAST > gpurun_out/r2_ncu_a.log 2>&1; echo "ncu tcx rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_nopiv_fused -s 4 -c 1 -o gpurun_out/r2_lu_fused -f $BENCH_FAST > gpurun_out/r2_ncu_b.log 2>&1; echo "ncu lu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_tc -s 4 -c 1 -o gpurun_out/r2_eval_tc -f $BENCH_FAST --eval-precision 1 > gpurun_out/r2_ncu_c.log 2>&1; echo "ncu tc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval64_mma -s 4 -c 1 -o gpurun_out/r2_eval64_mma -f $BENCH_FAST --eval-precision 2 > gpurun_out/r2_ncu_d.log 2>&1; echo "ncu eval64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_f32x2 -s 2 -c 1 -o gpurun_out/r2_eval_f32x2_c3g -f python profiles/tools/configs_probe.py --only C3g > gpurun_out/r2_ncu_e.log 2>&1; echo "ncu f32x2 rc=$?"
timeout 300 python profiles/tools/percook_probe.py > gpurun_out/r2_percook.jsonl 2>/dev/null; cat gpurun_out/r2_percook.jsonl
timeout 300 python profiles/tools/eval64_probe.py > gpurun_out/r2_eval64.jsonl 2>/dev/null; echo "eval64 rc=$?"
timeout 900 python tests/tools/accuracy_probe.py 256 1024 2048 4096 > gpurun_out/r2_accuracy.log 2>&1; cat gpurun_out/r2_accuracy.log
FD_LU_DEBUG=3 timeout 120 python profiles/tools/lu_step_probe.py 256 1024 > gpurun_out/r2_lu_probe.log 2>&1
tail -6 gpurun_out/r2_lu_probe.log
for n in 300 1500 4096 8192; do timeout 300 python profiles/tools/lu_determinism_probe.py $n 16; done > gpurun_out/r2_lu_determinism.log 2>&1; grep -E "mode count|^it" gpurun_out/r2_lu_determinism.log | head -12
(cd profiles/tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu && ./fp64_tput) > gpurun_out/r2_fp64_tput.log 2>&1; tail -4 gpurun_out/r2_fp64_tput.log
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_1gpu.json"))
print({k: j[k] for k in ("value", "ms_per_step", "phase_ms_last_step", "gpu_launches", "dtype")})
print(j["roofline"]["kernel"][:40], j["roofline"]["frac"], j["roofline"]["launch_ms"], j["config"]["eval_kernel"])
print("e2e", j.get("e2e")); print("factor", json.dumps(j.get("factor_ms_by_n")))
print("cpu", j.get("cpu_baseline"), j.get("cpu_baseline_1thread"))
for k, v in (j.get("other_configs") or {}).items(): print(k, v)
PY
