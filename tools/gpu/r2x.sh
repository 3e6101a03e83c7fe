mkdir -p gpurun_out
for v in "PGB_KBM_HSEQ=1" "PGB_KBM_SC32=1"; do
  echo "== $v"; env $v timeout 900 python -m pytest tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q -k "tensor_core or raw_entries or chopped_vs or high_order" 2>&1 | tail -n 3
done
L=gpurun_out/r2_kbm_flush_variants.log
for v in "PGB_KBM_FLUSH_AFTER=1" "PGB_KBM_HSEQ=1" "PGB_KBM_SC32=1"; do
  echo "# $v" >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
done
cat $L | cut -c1-330
