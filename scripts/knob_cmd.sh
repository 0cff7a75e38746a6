#!/bin/bash
# the measurement knob_variants.sh runs per library variant (edit per experiment)
python bench.py --no-cpu --only none | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['registration']; print(r['ms_per_step'], r['phasecorr_roofline']['kernel_ms'], r['phasecorr_roofline']['frac'], r['phasecorr_roofline_pow2']['kernel_ms'], r['phasecorr_roofline_pow2']['frac'])"
python scripts/prof_reg3d.py
