#!/bin/bash
mkdir -p gpurun_out
MVS_PC_SCALE_N2=1 timeout 600 python scripts/diag_reg.py > gpurun_out/diag_reg_n2.log 2>&1; tail -14 gpurun_out/diag_reg_n2.log | cut -c1-700
timeout 600 python scripts/diag_reg.py > gpurun_out/diag_reg.log 2>&1; tail -14 gpurun_out/diag_reg.log | cut -c1-700
timeout 900 python -m pytest tests/test_gpu_registration.py -m gpu -q > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_reg.log
grep -n "AssertionError\|passed\|failed" gpurun_out/pytest_reg.log | cut -c1-300
