#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 5000 --csv --log-file gpurun_out/launches_c3content.csv python scripts/bench_configs_big.py c3 > gpurun_out/c3c_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c3c_ncu.log
timeout 500 python scripts/bench_configs_big.py c3 > gpurun_out/big_c3.json 2> gpurun_out/big_c3.err; echo "c3 rc=$?"; cat gpurun_out/big_c3.json | tr -d '\n '; echo; tail -3 gpurun_out/big_c3.err
