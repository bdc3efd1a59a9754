#!/bin/bash
# round 2, GPU call H: panel-slab blocked sweeps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2h_pytest.log | tail -10
timeout 300 python profiles/tools/percook_probe.py > gpurun_out/r2h_percook.jsonl 2>/dev/null; cat gpurun_out/r2h_percook.jsonl
timeout 300 python profiles/tools/percook_probe.py > gpurun_out/r2h_percook2.jsonl 2>/dev/null; head -1 gpurun_out/r2h_percook2.jsonl
timeout 300 python profiles/tools/configs_probe.py --only C5s,C3g --v5 65536 > gpurun_out/r2h_configs.jsonl 2>&1; cat gpurun_out/r2h_configs.jsonl
