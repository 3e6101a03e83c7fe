set -x
mkdir -p gpurun_out
timeout 300 python tools/dropin_probe.py > gpurun_out/r2n_dropin.log 2>&1
cat gpurun_out/r2n_dropin.log | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_parity.py -x -q > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests.log
tail -n 12 gpurun_out/r2n_tests.log
