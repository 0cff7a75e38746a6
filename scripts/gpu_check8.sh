#!/bin/bash
mkdir -p gpurun_out
MVS_PC_DEBUG_STORE=1 timeout 600 python scripts/diag_reg2.py > gpurun_out/diag_reg2.log 2>&1; tail -6 gpurun_out/diag_reg2.log | cut -c1-600
timeout 600 python scripts/prof_c3.py > gpurun_out/c3.log 2>&1; tail -2 gpurun_out/c3.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-1500; tail -5 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuse_stencil -s 4 -c 1 -f -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_c2.log 2>&1; echo "ncu c2 rc=$?"
