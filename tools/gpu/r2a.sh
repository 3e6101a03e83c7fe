set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_config4.py -x -q -s > gpurun_out/r2a_config4.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_config4.log
python -m pytest tests -m gpu -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_gpu_tests.log
python tools/quick_time.py > gpurun_out/r2a_qt_raw.log 2>&1
QT_CHOP=1 python tools/quick_time.py > gpurun_out/r2a_qt_chop.log 2>&1
QT_CHOP=1 QT_SPARSE=1 python tools/quick_time.py > gpurun_out/r2a_qt_sparse.log 2>&1
PGB_KB_LEGACY=1 QT_CHOP=1 python tools/quick_time.py > gpurun_out/r2a_qt_legacy.log 2>&1
tail -3 gpurun_out/r2a_config4.log gpurun_out/r2a_gpu_tests.log; cat gpurun_out/r2a_qt_sparse.log
