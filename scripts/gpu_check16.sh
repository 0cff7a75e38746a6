#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_content.py -m gpu -q -x > gpurun_out/pytest_fus.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fus.log | cut -c1-300
timeout 500 python scripts/bench_configs_big.py c4 > gpurun_out/big_c4.json 2> gpurun_out/big_c4.err; echo "c4 rc=$?"
cat gpurun_out/big_c4.json | tr -d '\n '; echo; tail -3 gpurun_out/big_c4.err
