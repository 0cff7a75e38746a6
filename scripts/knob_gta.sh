#!/bin/bash
# content-weighted fusion per chunk for the libraries prebuilt under build_variants/
for v in build_variants/libmvs_gta*.so; do
  cp $v multiview_stitcher_b200/libmvs_b200.so
  echo "== $v"; python scripts/prof_content.py; python -m pytest tests/test_gpu_content.py -q -x 2>&1 | tail -1
done
