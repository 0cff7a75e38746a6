set -e
python -m pytest tests/test_gpu_fusion.py -m gpu -x -q 2>&1 | tail -2
python scripts/probe_r02c.py C3 C5row 2>&1 | tail -2
for v in "-DMVS_MAGIC_CVT=0"; do
  MVS_EXTRA_NVCC="$v" python -m multiview_stitcher_b200.build --force > /dev/null 2>&1
  echo "== $v"; python scripts/probe_r02c.py C3 C5row 2>&1 | tail -2
done
