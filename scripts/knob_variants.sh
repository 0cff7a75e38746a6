#!/bin/bash
cp multiview_stitcher_b200/libmvs_b200.so /tmp/default.so
for v in /tmp/default.so gpurun_variants/*.so; do
  cp $v multiview_stitcher_b200/libmvs_b200.so
  echo "== $v"
  python bench.py --no-cpu --only c3,c5 --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['configs']['C3']['fuse_blend_one_gpu']['ms'], d['configs']['C5']['ms'])"
done
cp /tmp/default.so multiview_stitcher_b200/libmvs_b200.so
