#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/diag_reg3d.py > gpurun_out/diag_reg3d.log 2>&1; tail -16 gpurun_out/diag_reg3d.log | cut -c1-900
