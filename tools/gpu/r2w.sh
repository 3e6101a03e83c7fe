mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q -k "ragged or tail or tensor_core or config4" 2>&1 | tail -n 3
PGB_KBM_FLUSH_AFTER=1 timeout 900 python -m pytest tests/test_gpu_config4.py -x -q -k "tensor_core or raw_entries" 2>&1 | tail -n 3
L=gpurun_out/r2_kbm_flush_variants.log
: > $L
for v in "" "PGB_KBM_FLUSH_AFTER=1"; do
  echo "# $v" >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
  env $v QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep branches32 >> $L
done
cat $L | cut -c1-400
RAGGED_TAIL=1 timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 1
