#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/diag_reg.py > gpurun_out/diag_reg.log 2>&1; cat gpurun_out/diag_reg.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuse_stencil -s 4 -c 1 -f -o gpurun_out/prof_c3 python scripts/prof_c3.py > gpurun_out/prof_c3.log 2>&1; echo "ncu c3 rc=$?"; tail -2 gpurun_out/prof_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ssim2d|fft_reg_pass|gauss_strip" -s 10 -c 10 -f -o gpurun_out/prof_reg python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_reg.log 2>&1; echo "ncu reg rc=$?"
