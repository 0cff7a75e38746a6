#!/bin/bash
# the measurement knob_variants.sh runs per library variant (edit per experiment)
python scripts/probe_e2e.py
