set -x
mkdir -p gpurun_out
QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py > gpurun_out/r2l_qt.log 2>&1
PGB_KA_REGS=84 QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py > gpurun_out/r2l_qt_ka84.log 2>&1
cat gpurun_out/r2l_qt.log gpurun_out/r2l_qt_ka84.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_tests.log
tail -n 6 gpurun_out/r2l_tests.log
