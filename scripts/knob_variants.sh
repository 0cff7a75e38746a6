#!/bin/bash
cp multiview_stitcher_b200/libmvs_b200.so /tmp/default.so
for v in gpurun_variants/*.so; do
  cp $v multiview_stitcher_b200/libmvs_b200.so
  echo "== $v"
  for i in 1 2; do python bench.py --no-cpu --only none --steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"; done
done
cp /tmp/default.so multiview_stitcher_b200/libmvs_b200.so
