#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fuse_kernel" -s 2 -c 1 -f -o gpurun_out/prof_c4 python scripts/bench_configs_big.py c4 > gpurun_out/prof_c4.log 2>&1; echo "ncu c4 rc=$?"; tail -2 gpurun_out/prof_c4.log | cut -c1-300
