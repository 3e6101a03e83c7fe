mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q -k "tensor_core or raw_entries or chopped_vs or high_order or sparse" 2>&1 | tail -n 3
L=gpurun_out/r2_kbm_flush_variants.log
for v in "PGB_KBM_NW8=1" "PGB_KBM_NW4=default"; do
  echo "# $v" >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
done
tail -n 6 $L | cut -c1-330
