set -x
mkdir -p gpurun_out
QT_CHOP=1 QT_SPARSE=1 python tools/quick_time.py 2>&1 | head -3 > gpurun_out/r2d_qt_rotsmem.log
PGB_KBM_ROT_REGS=1 QT_CHOP=1 QT_SPARSE=1 python tools/quick_time.py 2>&1 | head -3 > gpurun_out/r2d_qt_rotregs.log
python -m pytest tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q > gpurun_out/r2d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests.log
PGB_KBM_ROT_REGS=1 python -m pytest tests/test_gpu_config4.py -x -q > gpurun_out/r2d_tests_rotregs.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests_rotregs.log
python tools/ncu_fill.py > gpurun_out/r2d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:basis_mma -s 1 -c 1 -o gpurun_out/r2d_kbm python tools/ncu_fill.py > gpurun_out/r2d_ncu.log 2>&1
cat gpurun_out/r2d_qt_rotsmem.log gpurun_out/r2d_qt_rotregs.log; tail -n 4 gpurun_out/r2d_tests.log gpurun_out/r2d_tests_rotregs.log
