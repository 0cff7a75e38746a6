#!/bin/bash
# the measurement knob_variants.sh runs per library variant (edit per experiment)
python bench.py --no-cpu --only c4 | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']; print(c['C4']['fuse_blend']['ms'])"
