#!/bin/bash
# the measurement knob_variants.sh runs per library variant (edit per experiment)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ssim2d" -c 12 --csv --log-file gpurun_out/v.csv python scripts/prof_reg.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/v.csv | tail -1
