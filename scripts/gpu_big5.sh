#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_content.py -m gpu -q -x > gpurun_out/pytest_fus.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fus.log | cut -c1-300
for c in c3 c4; do
  timeout 500 python scripts/bench_configs_big.py $c > gpurun_out/big_$c.json 2> gpurun_out/big_$c.err; echo "$c rc=$?"
  cat gpurun_out/big_$c.json | tr -d '\n '; echo; tail -3 gpurun_out/big_$c.err
done
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "configs rc=$?"
cat gpurun_out/configs.json | tr -d '\n '; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_configs.csv python scripts/bench_configs.py > gpurun_out/c_ncu.log 2>&1; echo "ncu list configs rc=$?"
