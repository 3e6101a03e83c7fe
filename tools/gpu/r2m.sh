set -x
mkdir -p gpurun_out
timeout 300 python tools/dropin_probe.py > gpurun_out/r2m_dropin.log 2>&1
cat gpurun_out/r2m_dropin.log | cut -c1-420
QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py 2>&1 | cut -c1-330 | tee gpurun_out/r2m_qt.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log
tail -n 6 gpurun_out/r2m_tests.log
