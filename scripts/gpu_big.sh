#!/bin/bash
mkdir -p gpurun_out
for c in c3 c4; do
  timeout 500 python scripts/bench_configs_big.py $c > gpurun_out/big_$c.json 2> gpurun_out/big_$c.err; echo "$c rc=$?"
  cat gpurun_out/big_$c.json | tr -d '\n '; echo; tail -3 gpurun_out/big_$c.err
done
