#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_content.py tests/test_gpu_fusion.py -m gpu -q -x > gpurun_out/pytest_fus.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_fus.log | cut -c1-300
for c in c3 c4; do
  timeout 500 python scripts/bench_configs_big.py $c > gpurun_out/big_$c.json 2> gpurun_out/big_$c.err; echo "$c rc=$?"
  cat gpurun_out/big_$c.json | tr -d '\n '; echo; tail -3 gpurun_out/big_$c.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gauss_strip -s 6 -c 6 -f -o gpurun_out/prof_gauss python scripts/bench_configs.py > gpurun_out/prof_gauss.log 2>&1; echo "ncu gauss rc=$?"
