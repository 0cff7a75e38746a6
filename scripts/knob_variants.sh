#!/bin/bash
cp multiview_stitcher_b200/libmvs_b200.so /tmp/default.so
for v in /tmp/default.so gpurun_variants/*.so; do
  cp $v multiview_stitcher_b200/libmvs_b200.so
  echo "== $v"
  for i in 1 2; do python bench.py --no-cpu --only c3 --steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['configs']['C3']['fuse_blend_one_gpu']['ms'])"; done
done
cp /tmp/default.so multiview_stitcher_b200/libmvs_b200.so
