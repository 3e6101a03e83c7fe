set -x
mkdir -p gpurun_out
QT_CHOP=1 QT_SPARSE=1 python tools/quick_time.py > gpurun_out/r2c_qt_sparse.log 2>&1
python -m pytest tests/test_gpu_config4.py -x -q -s > gpurun_out/r2c_config4.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_config4.log
python -m pytest tests -m gpu -q > gpurun_out/r2c_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_gpu_tests.log
python tools/ncu_fill.py > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:basis_mma -s 1 -c 1 -o gpurun_out/r2c_kbm python tools/ncu_fill.py > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_qt_sparse.log; tail -n 5 gpurun_out/r2c_config4.log; tail -n 5 gpurun_out/r2c_gpu_tests.log
