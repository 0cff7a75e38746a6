#!/bin/bash
# round-2 validation: parity tests, smoke, bench (own arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r02b.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err ) 2>&1 | grep real; echo "bench rc=$?"
head -c 1500 gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
