#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_registration.py -m gpu -q -x > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_reg.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench.json')); r=b['registration']; print('pairs/s',r['pairs_per_sec'],'ms',r['ms_per_step'],'err',r['max_abs_shift_error_px'],'value',b['value'],'e2e',b['e2e']['value'])"
tail -3 gpurun_out/bench.err
