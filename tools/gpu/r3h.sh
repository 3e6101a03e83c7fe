mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_config4.py tests/test_gpu_parity.py tests/test_gpu_transfer.py -x -q 2>&1 | tail -n 2
L=gpurun_out/r2_kc_variants.log
QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep "branches32\|readme\|circle4" >> $L
QT_CHOP=1 QT_SPARSE=1 timeout 200 python tools/quick_time.py 2>&1 | grep "branches32\|readme\|circle4" >> $L
QT_CHOP=1 timeout 200 python tools/quick_time.py 2>&1 | grep "branches32" >> $L
cut -c1-330 $L
