#!/bin/bash
# A/B of compile-time knobs on one GPU box: build each variant HERE first, e.g.
#   MVS_EXTRA_NVCC="-DMVS_BX2=32" python -m multiview_stitcher_b200.build --force
#   cp multiview_stitcher_b200/libmvs_b200.so gpurun_variants/bx2_32.so      (gpurun_variants/ is git-ignored)
# then `gpurun -- bash scripts/knob_variants.sh "<command printing the numbers>"`.
CMD=${1:-"python bench.py --no-cpu --only none --steps 20 | python -c \"import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])\""}
cp multiview_stitcher_b200/libmvs_b200.so /tmp/default.so
for v in /tmp/default.so gpurun_variants/*.so; do
  cp $v multiview_stitcher_b200/libmvs_b200.so
  echo "== $v"
  eval "$CMD"
done
cp /tmp/default.so multiview_stitcher_b200/libmvs_b200.so
