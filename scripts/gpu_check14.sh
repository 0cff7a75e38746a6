#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "configs rc=$?"
cat gpurun_out/configs.json | tr -d '\n '; echo; tail -3 gpurun_out/configs.err
timeout 900 python scripts/diag_reg3d.py > gpurun_out/diag_reg3d.log 2>&1; grep -n "off ground\|quality" gpurun_out/diag_reg3d.log | tail -10 | cut -c1-200
