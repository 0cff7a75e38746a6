python scripts/probe_c4.py 2>&1 | grep -A1 "staged_order"
for v in "-DMVS_AFF_MINB=2" "-DMVS_AFF_MINB=5"; do
  MVS_EXTRA_NVCC="$v" python -m multiview_stitcher_b200.build --force > /dev/null 2>&1
  echo "== $v"; python scripts/probe_c4.py 2>&1 | grep -A1 "staged_order"
done
