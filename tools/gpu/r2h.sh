set -x
mkdir -p gpurun_out
timeout 300 python tools/dropin_probe.py > gpurun_out/r2h_dropin.log 2>&1
cat gpurun_out/r2h_dropin.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log
tail -n 5 gpurun_out/r2h_tests.log
