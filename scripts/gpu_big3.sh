#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/prof_host_c3.py > gpurun_out/prof_host_c3.log 2>&1; tail -70 gpurun_out/prof_host_c3.log | cut -c1-200
