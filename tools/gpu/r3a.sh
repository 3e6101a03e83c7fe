mkdir -p gpurun_out
PGB_KC_R8=1 timeout 900 python -m pytest tests/test_gpu_config4.py -x -q 2>&1 | tail -n 3
L=gpurun_out/r2_kc_variants.log
for v in "PGB_KC_R16=default" "PGB_KC_R8=1"; do
  echo "# $v" >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep "branches32\|readme" >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep "branches32\|readme" >> $L
done
tail -n 10 $L | cut -c1-330
