#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python scripts/diag_reg3d.py > gpurun_out/diag_reg3d.log 2>&1; grep -n "off ground\|quality" gpurun_out/diag_reg3d.log | tail -10 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "configs rc=$?"
cat gpurun_out/configs.json | tr -d '\n '; tail -5 gpurun_out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1; echo "ncu list rc=$?"
