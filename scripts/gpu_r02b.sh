#!/bin/bash
# round-2 validation: parity tests, smoke, bench (both arms), launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r02.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r02.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err ) 2>&1 | grep real; echo "bench rc=$?"
head -c 600 gpurun_out/bench_r02.json; echo; tail -5 gpurun_out/bench_r02.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err ) 2>&1 | grep real; echo "ref rc=$?"
head -c 900 gpurun_out/bench_r02_ref.json; echo
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --only none > gpurun_out/b_ncu.log 2>&1; echo "ncu list rc=$?"
