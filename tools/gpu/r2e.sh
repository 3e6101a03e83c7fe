set -x
mkdir -p gpurun_out
QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py 2>&1 | head -2 > gpurun_out/r2e_qt.log
timeout 300 python -m pytest tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q > gpurun_out/r2e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests.log
cat gpurun_out/r2e_qt.log; tail -n 4 gpurun_out/r2e_tests.log
